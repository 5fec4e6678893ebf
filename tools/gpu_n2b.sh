#!/bin/bash
# where does the table-sharded step lose time?  N=2
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export BENCH_HARD_LIMIT_S=120
stat() { python -c "import json,sys
try:
  j=json.loads(sys.stdin.read()); print('value %.1fM e2e %.1fM us/step %.2f'%(j['value']/1e6,j['e2e']['value']/1e6,j['ms_per_step']*1e3))
except Exception as e: print('n/a', e)"; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --cpu-seconds 0 --kernel-reps 3"
$TR > gpurun_out/n2b_default.log 2>&1; echo "sharded default: $(tail -n 1 gpurun_out/n2b_default.log | stat)"
$TR --replicate-mb 16 > gpurun_out/n2b_r16.log 2>&1; echo "sharded replicate<16MB: $(tail -n 1 gpurun_out/n2b_r16.log | stat)"
$TR --replicate-mb 128 > gpurun_out/n2b_r128.log 2>&1; echo "sharded replicate<128MB: $(tail -n 1 gpurun_out/n2b_r128.log | stat)"
FR_SHARD_NOWAIT=1 $TR > gpurun_out/n2b_nowait.log 2>&1; echo "sharded nowait (racy, timing only): $(tail -n 1 gpurun_out/n2b_nowait.log | tail -c 400)"
$TR --shard replicated > gpurun_out/n2b_repl.log 2>&1; echo "replicated: $(tail -n 1 gpurun_out/n2b_repl.log | stat)"
$TR --streams 16 > gpurun_out/n2b_s16.log 2>&1; echo "sharded s16: $(tail -n 1 gpurun_out/n2b_s16.log | stat)"
