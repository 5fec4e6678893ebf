#!/usr/bin/env python3
"""One GPU, world = 1: the same batches through fr_infer and through the sharded entry point (fr_shard_infer on a
one-rank 'shard': no peer, nothing exchanged).  Isolates what the sharded step costs beyond the exchange itself.

  python tools/shard_gap.py [steps] [workers]
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "gpu-fpga-recommendation-system_b200"))
import fleetrec  # noqa: E402
from fleetrec import catalogue  # noqa: E402
from oracle import oracle  # noqa: E402  (index / weight generators only)

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
nw = int(sys.argv[2]) if len(sys.argv) > 2 else 12
B = 2048
cat = catalogue.load("small")
dims = cat.layer_dims
W, b = oracle.make_weights(dims, seed=42)
pool = [torch.from_numpy(oracle.zipf_indices(cat, B, seed=10 + i)).cuda() for i in range(16)]


def run(mode):
    eng = fleetrec.Engine(cat, max_batch=B)
    if mode != "infer":
        owner = [0] * cat.n_tables if mode == "shard_owned" else [-1] * cat.n_tables
        eng.shard_init(0, 1, owner)
    eng.fill_hash(seed=1)
    eng.load_mlp(W, b)
    if mode != "infer":
        eng.shard_attach_local([eng])
    ws = [fleetrec.Worker(eng) for _ in range(nw)]
    sc = [torch.empty(B, dtype=torch.float32, device="cuda") for _ in range(nw)]

    def step(i):
        w = i % nw
        if mode == "infer":
            eng.infer_async(pool[i % 16].data_ptr(), sc[w].data_ptr(), B, ws[w])
        else:
            eng.shard_infer(pool[i % 16].data_ptr(), B, sc[w].data_ptr(), ws[w])
    for i in range(16 * nw * 2):
        step(i)
    for w in ws:
        eng.sync(w)
    t0 = time.perf_counter()
    for i in range(steps):
        step(i)
    t1 = time.perf_counter()
    for w in ws:
        eng.sync(w)
    t2 = time.perf_counter()
    print(f"{mode:12s}: {(t2 - t0) / steps * 1e6:6.2f} us per step (host enqueue {(t1 - t0) / steps * 1e6:.2f}), "
          f"{eng.launch_count()} launches", flush=True)
    for w in ws:
        w.close()
    eng.close()


for m in ("infer", "shard_owned", "shard_repl", "infer"):
    run(m)
