#!/usr/bin/env python3
"""Phase timeline of the one-launch MLP chain (tc_mlp_chain_kernel), CTA 0, first item tiles.

  FR_CHAIN_PROF=1 python tools/chain_timeline.py [model] [batch] [clusters]

Prints, per item-tile iteration and phase, microseconds since the first stamp:
  mma: TMEM free / first smem slot full / last MMA issued     (MMA issuer thread)
  epi: accumulator complete seen / TMEM handed back / this warp's stores complete   (epilogue warp 2)
  prod: first / last wait for the previous layer's chunk     (TMA producer)
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "gpu-fpga-recommendation-system_b200"))
os.environ["FR_CHAIN_PROF"] = "1"
if len(sys.argv) > 3:
    os.environ["FR_TC_MAX_CLUSTERS"] = sys.argv[3]

import fleetrec  # noqa: E402
from fleetrec import _capi, catalogue  # noqa: E402

model = sys.argv[1] if len(sys.argv) > 1 else "small"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
cat = catalogue.load(model).with_row_cap(64)
dims = cat.layer_dims
rng = np.random.default_rng(0)
W = [(rng.standard_normal((dims[k], dims[k + 1])) / np.sqrt(dims[k])).astype(np.float32) for k in range(4)]
b = [np.zeros(dims[k + 1], np.float32) for k in range(4)]
eng = fleetrec.Engine(cat, max_batch=B)
eng.load_mlp(W, b)
x = rng.uniform(-1, 1, (B, dims[0])).astype(np.float32)
for _ in range(3):
    eng.mlp_only(x)
raw = C.CDLL(_capi.LIB_PATH)
raw.frdbg_chain_timeline.argtypes = [C.c_void_p, C.POINTER(C.c_longlong), C.c_int]
buf = (C.c_longlong * 256)()
n = raw.frdbg_chain_timeline(eng._h, buf, 256)
assert n == 256, "no timeline (FR_CHAIN_PROF / chain kernel not used?)"
t = np.array(buf[:], dtype=np.int64).reshape(4, 8, 8)
probe = t[3, 7, :3].copy()
t[3, 7, :] = 0
brk = t[0, 4:8, :5].copy()      # iteration 0, rows 4..7: epilogue breakdown of phases 0..3 (cycles, warp 2)
t[0, 4:8, :] = 0
t0 = t[t > 0].min()
us = lambda v: "   -  " if v == 0 else f"{(v - t0) / 1965.0:6.1f}"
print(f"{model} B={B} dims={dims}: us since first stamp (1965 MHz)")
print("it ph | mma: tmem_free first_full last_issue | epi: full_seen tmem_released stores_done | prod: ready_wait ready_ok")
for it in range(4):
    for ph in range(8):
        r = t[it, ph]
        if not r.any():
            continue
        print(f"{it:2d} {ph:2d} |      {us(r[0])}   {us(r[1])}    {us(r[2])}   |      {us(r[3])}    {us(r[4])}      {us(r[5])}    |       {us(r[6])}   {us(r[7])}")
print("epilogue warp 2, iteration 0, us per phase: tmem_ld+wait | wait_read<1> | bias/relu/round/st.shared | store issue | fence.proxy.async+syncwarp")
for ph in range(4):
    if brk[ph].any():
        print(f"   ph {ph}: " + "  ".join(f"{v / 1965.0:6.2f}" for v in brk[ph]))
if probe[0]:
    print(f"probe warps of CTA 0: {probe[0]} iterations x 16 KB in {probe[1] / 1965.0:.1f} us = "
          f"{probe[0] * 16384 / (probe[1] / 1.965):.1f} GB/s through the LSU")
eng.close()
