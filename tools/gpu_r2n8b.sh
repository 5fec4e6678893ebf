#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export BENCH_HARD_LIMIT_S=200
N=${N:-8}
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/n${N}_default.log 2>&1
echo "exit $?"; grep '^{' gpurun_out/n${N}_default.log | tail -n1 | cut -c1-300; grep -o '"e2e": {[^}]*}' gpurun_out/n${N}_default.log | tail -n1
