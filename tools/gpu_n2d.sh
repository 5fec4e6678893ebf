#!/bin/bash
# 2-GPU check of the sharded path: shard tests (incl. column-sliced step), then table-sharded and replicated bench
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export BENCH_HARD_LIMIT_S=200
stat() { python -c "import json,sys
try:
  j=json.loads(sys.stdin.read()); print('value %.1fM e2e %.1fM us/step %.2f e2e us/step %.2f launches %d h2d %d'%(j['value']/1e6,j['e2e']['value']/1e6,j['ms_per_step']*1e3,j['e2e']['ms_per_step']*1e3,j['gpu_launches'],j['e2e']['h2d_bytes_per_step']))
except Exception as e: print('n/a', e)"; }
timeout 300 python -m pytest tests/test_shard.py -q -m gpu 2>&1 | tail -n 3
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --cpu-seconds 0 --kernel-reps 3"
timeout 300 $TR > gpurun_out/n2_sharded_r01.log 2>&1; echo "sharded: $(tail -n 1 gpurun_out/n2_sharded_r01.log | stat)"; tail -n 3 gpurun_out/n2_sharded_r01.log | cut -c1-300 | grep -v "^{"
timeout 300 $TR --shard replicated > gpurun_out/n2_replicated_r01.log 2>&1; echo "replicated: $(tail -n 1 gpurun_out/n2_replicated_r01.log | stat)"
