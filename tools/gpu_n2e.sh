#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export BENCH_HARD_LIMIT_S=200
stat() { python -c "import json,sys
try:
  j=json.loads(sys.stdin.read()); print('value %.1fM e2e %.1fM us/step %.2f e2e us/step %.2f'%(j['value']/1e6,j['e2e']['value']/1e6,j['ms_per_step']*1e3,j['e2e']['ms_per_step']*1e3))
except Exception as e: print('n/a', e)"; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --cpu-seconds 0 --kernel-reps 3"
FR_SHARD_NOWAIT=2 FR_SHARD_PRIVATE=1 timeout 300 $TR --replicate-mb 100000 > gpurun_out/n2_z1.log 2>&1; echo "all replicated, no flags, concat in a PRIVATE buffer: $(tail -n 1 gpurun_out/n2_z1.log | stat)"
FR_SHARD_NOWAIT=2 timeout 300 $TR --replicate-mb 100000 > gpurun_out/n2_z2.log 2>&1; echo "all replicated, no flags, concat in the exchange region: $(tail -n 1 gpurun_out/n2_z2.log | stat)"
