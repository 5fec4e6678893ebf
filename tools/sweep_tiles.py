"""Per-kernel device times for every tcgen05 tile configuration (run on the GPU box)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "gpu-fpga-recommendation-system_b200"))
import fleetrec
from fleetrec import catalogue
from oracle import oracle
model = sys.argv[1] if len(sys.argv) > 1 else "small"
cat = catalogue.load(model).with_row_cap(200000)
dims = cat.layer_dims
W, b = oracle.make_weights(dims)
for tiles in ("128,128,256,1", "256,256,256,1", "128,128,256,2", "256,128,256,2", "256,256,256,2"):
    os.environ["FR_TC_TILES"] = tiles
    eng = fleetrec.Engine(cat, max_batch=16384)
    eng.fill_hash()
    eng.load_mlp(W, b)
    for B in (2048, 16384):
        idx = oracle.zipf_indices(cat, B)
        ms = eng.time_kernels(idx, B, reps=50)
        fl = [0, 2. * B * dims[0] * dims[1], 2. * B * dims[1] * dims[2], 2. * B * dims[2] * dims[3]]
        print(f"{model} tiles={tiles:16s} B={B:6d} gather {ms[0]*1e3:7.1f}us | " +
              " | ".join(f"L{k} {ms[k]*1e3:7.1f}us {fl[k]/ms[k]/1e9:6.0f}TF" for k in (1, 2, 3)) +
              f" | sum {sum(ms[:4])*1e3:7.1f}us -> {B/sum(ms[:4])/1e3:7.2f} Minf/s serial")
    eng.close()
