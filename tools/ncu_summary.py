#!/usr/bin/env python3
"""Turn the ncu artefacts of one gpurun call (gpurun_out/) into the tracked summaries under profiles/.

  python tools/ncu_summary.py r01            # reads gpurun_out/launches_r01.csv, prof_r01_*.ncu-rep
Writes profiles/<round>_launches.csv (per-kernel launch count / avg time / share of the step),
profiles/<round>_<capture>.csv (key raw metrics per captured launch) and profiles/<round>_bench.json.
"""
import collections
import csv
import glob
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum",
        "sm__cycles_elapsed.max", "sm__cycles_active.avg", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor"]


def launches(tag):
    src = os.path.join(OUT, f"launches_{tag}.csv")
    if not os.path.exists(src):
        return
    rows = [r for r in csv.reader(open(src, errors="replace")) if len(r) > 10]
    hdr = rows[0]
    ki, vi, gi, bi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Block Size")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        agg.setdefault((r[ki], r[gi], r[bi]), []).append(v)
    tot = sum(sum(v) for v in agg.values())
    with open(os.path.join(PROF, f"{tag}_launches.csv"), "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "grid", "block", "launches", "avg_us", "min_us", "max_us", "share_of_gpu_time"])
        for (k, g, b), v in agg.items():
            w.writerow([k, g, b, len(v), f"{sum(v) / len(v) / 1e3:.3f}", f"{min(v) / 1e3:.3f}", f"{max(v) / 1e3:.3f}",
                        f"{sum(v) / tot:.4f}"])
    print(f"profiles/{tag}_launches.csv: {len(agg)} kernels, {sum(len(v) for v in agg.values())} launches")


def capture(path, tag):
    name = os.path.basename(path)[:-len(".ncu-rep")]
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    if len(rows) < 3:
        return
    hdr, units = rows[0], rows[1]
    cols = [hdr.index("Kernel Name")] + [hdr.index(k) for k in KEYS if k in hdr]
    with open(os.path.join(PROF, f"{name.replace('prof_', '')}.csv"), "w", newline="") as f:
        w = csv.writer(f)
        w.writerow([hdr[c] + (f" [{units[c]}]" if units[c] else "") for c in cols])
        for r in rows[2:]:
            w.writerow([r[c][:140] for c in cols])
    print(f"profiles/{name.replace('prof_', '')}.csv: {len(rows) - 2} launches")


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
    os.makedirs(PROF, exist_ok=True)
    launches(tag)
    for p in sorted(glob.glob(os.path.join(OUT, f"prof_{tag}_*.ncu-rep"))):
        capture(p, tag)
    b = os.path.join(OUT, "bench.log")
    if os.path.exists(b):
        lines = [l for l in open(b) if l.startswith("{")]
        if lines:
            open(os.path.join(PROF, f"{tag}_bench.json"), "w").write(lines[-1])


if __name__ == "__main__":
    main()
