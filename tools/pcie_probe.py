#!/usr/bin/env python3
"""Host<->device copy rates of this box: what bounds bench.py's e2e leg (385 KB of indices per step)."""
import torch
torch.cuda.init()
def rate(nbytes, reps, streams=1, d2h=False):
    hs = [torch.empty(nbytes, dtype=torch.uint8).pin_memory() for _ in range(streams)]
    ds = [torch.empty(nbytes, dtype=torch.uint8, device="cuda") for _ in range(streams)]
    ss = [torch.cuda.Stream() for _ in range(streams)]
    def go(n):
        for i in range(n):
            k = i % streams
            with torch.cuda.stream(ss[k]):
                (hs[k].copy_(ds[k], non_blocking=True) if d2h else ds[k].copy_(hs[k], non_blocking=True))
    go(streams * 3); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in ss: s.wait_event(e0)
    go(reps)
    for s in ss:
        ev = torch.cuda.Event(); ev.record(s); torch.cuda.current_stream().wait_event(ev)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    return nbytes * reps / ms / 1e6, ms / reps * 1e3
for nb, reps in ((256 << 20, 10), (4 << 20, 200), (385024, 2000), (65536, 2000), (8192, 2000)):
    for st in (1, 4, 12):
        g, us = rate(nb, reps, st)
        print(f"H2D {nb:>10} B x{reps} on {st:2d} streams: {g:6.1f} GB/s  {us:8.2f} us/copy")
g, us = rate(8192, 2000, 4, d2h=True); print(f"D2H 8192 B: {g:.2f} GB/s {us:.2f} us/copy")
g, us = rate(256 << 20, 10, 1, d2h=True); print(f"D2H 256 MB: {g:.1f} GB/s")
