#!/usr/bin/env python3
"""Host<->device copy rates of this box: what bounds bench.py's e2e leg (385 KB of indices per step).

Copies are captured into a CUDA graph (64 per stream per replay) so the host's per-call cost (~11 us per
torch copy_) is not what is measured: the rows are what the copy engine itself sustains."""
import torch

torch.cuda.init()


def rate(nbytes, streams=1, d2h=False, per_graph=64, replays=10):
    hs = [torch.empty(nbytes, dtype=torch.uint8).pin_memory() for _ in range(streams)]
    ds = [torch.empty(nbytes, dtype=torch.uint8, device="cuda") for _ in range(streams)]
    side = [torch.cuda.Stream() for _ in range(streams)]
    cap = torch.cuda.Stream()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(cap):
        for k in range(streams):            # warm the copy path outside capture
            ds[k].copy_(hs[k], non_blocking=True)
        cap.synchronize()
        with torch.cuda.graph(g, stream=cap):
            for k in range(streams):
                side[k].wait_stream(cap)
                with torch.cuda.stream(side[k]):
                    for _ in range(per_graph):
                        (hs[k].copy_(ds[k], non_blocking=True) if d2h else ds[k].copy_(hs[k], non_blocking=True))
            for k in range(streams):
                cap.wait_stream(side[k])
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(replays):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    n = replays * per_graph * streams
    ms = e0.elapsed_time(e1)
    return nbytes * n / ms / 1e6, ms / n * 1e3


for nb in (256 << 20, 4 << 20, 385024, 65536, 8192):
    for st in (1, 2, 4, 12):
        pg = 2 if nb > (64 << 20) else 64
        gbs, us = rate(nb, st, per_graph=pg, replays=3 if nb > (64 << 20) else 10)
        print(f"H2D {nb:>10} B, {st:2d} streams (graph replay): {gbs:6.1f} GB/s  {us:8.2f} us per copy (aggregate)")
gbs, us = rate(8192, 4, d2h=True)
print(f"D2H       8192 B,  4 streams: {gbs:6.2f} GB/s {us:8.2f} us per copy")
gbs, us = rate(256 << 20, 1, d2h=True, per_graph=2, replays=3)
print(f"D2H  268435456 B,  1 stream : {gbs:6.1f} GB/s")
