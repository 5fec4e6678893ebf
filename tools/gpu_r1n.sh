#!/bin/bash
# reduced-precision tables, TCP ingest, latency-mode tiles
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export BENCH_HARD_LIMIT_S=400
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "reduced_precision or ingest or batcher or fused or end_to_end" 2>&1 | tail -n 6
timeout 600 python -m pytest tests -q -m gpu 2>&1 | tail -n 3
timeout 400 python bench.py --workload sweep --sweep-launches 500 > gpurun_out/sweep2.log 2>&1; tail -n 1 gpurun_out/sweep2.log | python -c "
import json,sys
j=json.loads(sys.stdin.read())
for r in j['sweep']: print(r['batch'], round(r['dev_p50_us'],1), round(r['dev_p99_us'],1), round(r['host_p50_us'],1), round(r['inferences_per_s']/1e6,2))"
timeout 600 python bench.py --workload stress --table-dtype f16 --stress-rows 10000000 > gpurun_out/stress_f16.log 2>&1; tail -n 1 gpurun_out/stress_f16.log | cut -c1-1500
timeout 300 python bench.py --workload stress --table-dtype bf16 --stress-rows 2000000 > gpurun_out/stress_bf16.log 2>&1; tail -n 1 gpurun_out/stress_bf16.log | cut -c1-700
