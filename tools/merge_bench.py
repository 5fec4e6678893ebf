"""What do Cartesian-merged tables (fleetrec/merge.py) do to the small model's lookup?  (GPU box)

The small model's lookup is transaction-bound, not HBM-bound: 47 random accesses per item, most of them 16- and 32-byte
rows (70 % of the HBM roofline at 16384 items, 22 % at 2048).  Merging table pairs removes one
access per pair and widens the rows.  For a few byte budgets: plan, build the merged images on the device, check the
concat vectors bit for bit against the un-merged engine, time the lookup alone.  Prints one JSON line."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "gpu-fpga-recommendation-system_b200"))
import torch  # noqa: E402

import fleetrec  # noqa: E402
from fleetrec import catalogue, merge  # noqa: E402
from oracle import oracle  # noqa: E402

cat = catalogue.load("small")
tables = oracle.make_tables(cat, "hash", seed=0x5EED)
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
out = []
ref = {}
for budget_mb in (0, 256, 4096, 65536):
    plan = merge.plan_merges(cat, budget_mb << 20)
    mm = merge.apply_merges(cat, plan.pairs)
    eng = merge.build(mm, tables, max_batch=16384)
    w = fleetrec.Worker(eng)
    row = dict(budget_mb=budget_mb, pairs=len(plan.pairs), lookups_per_item=merge.lookups_per_item(mm),
               extra_gb=plan.extra_bytes / 1e9, table_gb=eng.table_bytes() / 1e9)
    for B in (2048, 16384):
        idx = oracle.uniform_indices(cat, B, seed=4321)
        midx = torch.from_numpy(mm.remap(idx)).cuda()
        got = torch.empty(B, cat.concat_floats, dtype=torch.float32, device="cuda")
        eng.gather_only_async(midx, got, B, w)
        eng.sync(w)
        g = got.cpu().numpy()
        if budget_mb == 0:
            ref[B] = g
            assert np.array_equal(g[:64].view(np.uint32), oracle.gather_hashed(cat, 0x5EED, idx[:64]).view(np.uint32))
        else:
            assert np.array_equal(g.view(np.uint32), ref[B].view(np.uint32)), "merged lookup differs from the un-merged one"
        for _ in range(5):
            eng.gather_only_async(midx, got, B, w)
        eng.sync(w)
        eng.mark(0, w)
        for _ in range(50):
            eng.gather_only_async(midx, got, B, w)
        eng.mark(1, w)
        ms = eng.elapsed_ms(w) / 50
        alg = B * (sum(t.dim for t in cat.tables) * 4 + row["lookups_per_item"] * 4 + cat.concat_floats * 4)
        row[f"B{B}"] = dict(us=ms * 1e3, gbs=alg / (ms * 1e-3) / 1e9, frac_of_hbm_peak=alg / (ms * 1e-3) / 1e9 / peak)
    w.close()
    eng.close()
    torch.cuda.empty_cache()
    out.append(row)
print(json.dumps({"what": "small model lookup alone, uniform indices, Cartesian-merged tables under a byte budget",
                  "hbm_peak_gbs": peak, "rows": out}))
