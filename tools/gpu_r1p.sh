#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export BENCH_HARD_LIMIT_S=300
summ='import json,sys
j=json.loads(sys.stdin.read()); print("value %.1f e2e %.1f M/s  step %.2f us e2e %.2f us" % (j["value"]/1e6, j["e2e"]["value"]/1e6, j["ms_per_step"]*1e3, j["e2e"]["ms_per_step"]*1e3), j["host_enqueue_us_per_step"])'
for z in 2 0; do
  echo "=== bench FR_ZEROCOPY=$z"
  FR_ZEROCOPY=$z timeout 400 python bench.py --cpu-seconds 0 --gather-batch 2048 > gpurun_out/bench_tmp.log 2>&1; echo "exit $?"; tail -n 1 gpurun_out/bench_tmp.log | python -c "$summ"
done
timeout 200 python tools/pcie_probe.py > gpurun_out/pcie_r01.log 2>&1; tail -n 3 gpurun_out/pcie_r01.log
