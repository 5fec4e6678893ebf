#!/bin/bash
# hardware work queues: 12 worker streams on the default 8 connections share queues (false dependencies; a spinning
# flag kernel at the head of a queue holds back another worker's kernels behind it)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export BENCH_HARD_LIMIT_S=200
stat() { python -c "import json,sys
try:
  j=json.loads(sys.stdin.read()); print('value %.1fM e2e %.1fM us/step %.2f e2e us/step %.2f host %s'%(j['value']/1e6,j['e2e']['value']/1e6,j['ms_per_step']*1e3,j['e2e']['ms_per_step']*1e3, j['host_enqueue_us_per_step']))
except Exception as e: print('n/a', e)"; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --cpu-seconds 0 --kernel-reps 3"
for mc in 8 32; do
  for rep in 1 2; do
    CUDA_DEVICE_MAX_CONNECTIONS=$mc timeout 300 $TR > gpurun_out/n2_mc${mc}_$rep.log 2>&1; echo "sharded, $mc connections, run $rep: $(tail -n 1 gpurun_out/n2_mc${mc}_$rep.log | stat)"
  done
done
CUDA_DEVICE_MAX_CONNECTIONS=32 timeout 300 python bench.py --cpu-seconds 0 --gather-batch 2048 > gpurun_out/n1_mc32.log 2>&1; echo "1 GPU, 32 connections: $(tail -n 1 gpurun_out/n1_mc32.log | stat)"
CUDA_DEVICE_MAX_CONNECTIONS=8 timeout 300 python bench.py --cpu-seconds 0 --gather-batch 2048 > gpurun_out/n1_mc8.log 2>&1; echo "1 GPU, 8 connections: $(tail -n 1 gpurun_out/n1_mc8.log | stat)"
