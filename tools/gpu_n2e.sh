#!/bin/bash
# where the table-sharded step loses against replicated tables (2 GPUs)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export BENCH_HARD_LIMIT_S=200
stat() { python -c "import json,sys
try:
  j=json.loads(sys.stdin.read()); print('value %.1fM e2e %.1fM us/step %.2f e2e us/step %.2f host %s'%(j['value']/1e6,j['e2e']['value']/1e6,j['ms_per_step']*1e3,j['e2e']['ms_per_step']*1e3, j['host_enqueue_us_per_step']))
except Exception as e: print('n/a', e)"; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --cpu-seconds 0 --kernel-reps 3"
FR_SHARD_NOWAIT=2 timeout 300 $TR --replicate-mb 100000 > gpurun_out/n2_x1.log 2>&1; echo "all replicated, NO flag kernel (timing only): $(tail -n 1 gpurun_out/n2_x1.log | stat)"
FR_SHARD_NOWAIT=2 timeout 300 $TR > gpurun_out/n2_x2.log 2>&1; echo "sharded push, NO flag kernel (racy, timing only): $(tail -n 1 gpurun_out/n2_x2.log | stat)"
FR_SHARD_SLOTS=24 timeout 300 $TR --streams 24 > gpurun_out/n2_x3.log 2>&1; echo "sharded, 24 workers / slots: $(tail -n 1 gpurun_out/n2_x3.log | stat)"; grep -i "error" gpurun_out/n2_x3.log | head -3
