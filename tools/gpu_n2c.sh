#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export BENCH_HARD_LIMIT_S=150
stat() { python -c "import json,sys
try:
  j=json.loads(sys.stdin.read()); print('value %.1fM e2e %.1fM us/step %.2f launches %d'%(j['value']/1e6,j['e2e']['value']/1e6,j['ms_per_step']*1e3,j['gpu_launches']))
except Exception as e: print('n/a', e)"; }
FR_SHARD_ONE_LAUNCH=1 timeout 300 python -m pytest tests/test_shard.py -q -m gpu 2>&1 | tail -n 2; timeout 300 python -m pytest tests/test_shard.py -q -m gpu 2>&1 | tail -n 2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --cpu-seconds 0 --kernel-reps 3"
FR_SHARD_ONE_LAUNCH=1 $TR > gpurun_out/n2c_sharded.log 2>&1; echo "sharded one-launch: $(tail -n 1 gpurun_out/n2c_sharded.log | stat)"
$TR > gpurun_out/n2c_sharded2.log 2>&1; echo "sharded 3-kernel: $(tail -n 1 gpurun_out/n2c_sharded2.log | stat)"
$TR --shard replicated > gpurun_out/n2c_repl.log 2>&1; echo "replicated: $(tail -n 1 gpurun_out/n2c_repl.log | stat)"
