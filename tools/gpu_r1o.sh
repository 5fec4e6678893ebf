#!/bin/bash
# one-launch MLP chain: parity first, then the bench with and without it, then the whole GPU suite
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export BENCH_HARD_LIMIT_S=300
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "chain_kernel" 2>&1 | tail -n 15
echo "=== bench (chain)"
timeout 400 python bench.py --cpu-seconds 0 > gpurun_out/bench_chain.log 2>&1; echo "exit $?"; tail -n 1 gpurun_out/bench_chain.log | cut -c1-3000
echo "=== bench (FR_CHAIN=0)"
FR_CHAIN=0 timeout 400 python bench.py --cpu-seconds 0 > gpurun_out/bench_nochain.log 2>&1; echo "exit $?"; tail -n 1 gpurun_out/bench_nochain.log | cut -c1-600
echo "=== all gpu tests"
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -n 5
