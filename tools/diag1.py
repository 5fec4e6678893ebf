import sys, os, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "gpu-fpga-recommendation-system_b200"))
import fleetrec
from fleetrec import catalogue
np.set_printoptions(linewidth=200, suppress=True)
for model, B in (("medium", 130), ("small", 32), ("small", 130)):
    cat = catalogue.load(model).with_row_cap(64)
    dims = cat.layer_dims
    eng = fleetrec.Engine(cat, mlp_mode=fleetrec.FR_MLP_LINEAR, max_batch=256)
    eng.load_mlp([np.ones((dims[k], dims[k + 1]), np.float32) for k in range(4)])
    x = np.ones((B, dims[0]), np.float32)
    h1 = eng.layer_only(0, x, dims[1])
    print(model, B, "h1 unique", np.unique(h1)[:10], "rows differ?", np.ptp(h1, axis=0).max(), "cols 0..40", h1[0, :40])
    print("  h1 col pattern by n%8:", [float(h1[0, n::8].mean()) for n in range(8)], " by row m: ", h1[:, 0][:16], h1[-2:, 0])
    h2 = eng.layer_only(1, np.full((B, dims[1]), float(dims[0]), np.float32), dims[2])
    print("  h2 unique", np.unique(h2)[:10] / dims[0])
    s = eng.layer_only(2, np.full((B, dims[2]), 1.0, np.float32), 1)
    print("  s unique", np.unique(s)[:10])
    # mixed rows like the chain KAT
    xm = np.ones((B, dims[0]), np.float32); xm[::2] = 0
    h1 = eng.layer_only(0, xm, dims[1])
    print("  mixed rows: h1[:6,0]", h1[:6, 0], "unique", np.unique(h1)[:8])
    out = eng.mlp_only(xm)
    print("  mixed rows mlp_only unique:", np.unique(out))
    out = eng.mlp_only(x)
    print("  ones mlp_only unique:", np.unique(out))
    eng.close()
