#!/bin/bash
# 2-GPU call: sharded device tests, bench N=2 (table-sharded vs replicated), config 4 (large model sharded)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export BENCH_TRACE=1
nvidia-smi -L
nvidia-smi topo -m | head -6
run() { name=$1; shift; echo "=== $name"; timeout "$@" > gpurun_out/$name.log 2>&1; echo "exit $? ($name)"; grep "^\[bench" gpurun_out/$name.log | tail -n 1; tail -n ${TAILN:-1} gpurun_out/$name.log | cut -c1-${CUT:-900}; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
TAILN=8 run t_shard 300 python -m pytest tests/test_shard.py -q -m gpu
run n2_sharded 240 $TR bench.py --gpus 2 --cpu-seconds 0 --kernel-reps 5
run n2_replicated 240 $TR bench.py --gpus 2 --cpu-seconds 0 --kernel-reps 5 --shard replicated
run n2_sharded_s4 240 $TR bench.py --gpus 2 --cpu-seconds 0 --kernel-reps 5 --streams 4
FR_GRAPHS=0 run n2_sharded_nograph 240 $TR bench.py --gpus 2 --cpu-seconds 0 --kernel-reps 5
CUT=1500 run n2_large_b4096 400 $TR bench.py --gpus 2 --model large --batch 2048 --steps 500 --cpu-seconds 0 --kernel-reps 5
CUT=1500 run n2_large_b16384 400 $TR bench.py --gpus 2 --model large --batch 8192 --steps 200 --cpu-seconds 0 --kernel-reps 5
run n1_default 240 python bench.py --cpu-seconds 0 --kernel-reps 5
