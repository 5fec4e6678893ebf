"""Launch every kernel of the step a few times at one batch size (ncu target; also prints times)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "gpu-fpga-recommendation-system_b200"))
import fleetrec
from fleetrec import catalogue
from oracle import oracle
model = sys.argv[1] if len(sys.argv) > 1 else "small"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
cat = catalogue.load(model).with_row_cap(2000000)
dims = cat.layer_dims
W, b = oracle.make_weights(dims)
eng = fleetrec.Engine(cat, max_batch=B)
eng.fill_hash()
eng.load_mlp(W, b)
idx = oracle.uniform_indices(cat, B, seed=4321)
ms = eng.time_kernels(idx, B, reps=reps)
print(model, B, os.environ.get("FR_TC_TILES", "default"), " ".join(f"{m*1e3:.1f}us" for m in ms))
eng.close()
