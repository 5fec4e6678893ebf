#!/bin/bash
# 4-GPU check of the table-sharded bench (column-sliced index feed)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export BENCH_HARD_LIMIT_S=200
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --cpu-seconds 0 --kernel-reps 3 > gpurun_out/n4_sharded_r01.log 2>&1
echo "exit $?"; tail -n 1 gpurun_out/n4_sharded_r01.log | python -c "
import json,sys
try:
  j=json.loads(sys.stdin.read()); print('value %.1fM e2e %.1fM us/step %.2f e2e us/step %.2f h2d %d (max per rank %d)'%(j['value']/1e6,j['e2e']['value']/1e6,j['ms_per_step']*1e3,j['e2e']['ms_per_step']*1e3,j['e2e']['h2d_bytes_per_step'],j['e2e']['h2d_bytes_per_rank_max']))
except Exception as e: print('n/a', e)"
grep -i "error\|assert" gpurun_out/n4_sharded_r01.log | head -5
