#!/bin/bash
# 8-GPU validation: table-sharded (default) and replicated at N=8, sharded at N=4
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export BENCH_TRACE=1 BENCH_HARD_LIMIT_S=150
run() { name=$1; shift; echo "=== $name"; "$@" > gpurun_out/$name.log 2>&1; echo "exit $? ($name)"; grep "^\[bench" gpurun_out/$name.log | tail -n 1; tail -n 1 gpurun_out/$name.log | cut -c1-${CUT:-330}; }
TR() { n=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --cpu-seconds 0 --kernel-reps 5 "$@"; }
run n8_sharded TR 8
run n8_replicated TR 8 --shard replicated
run n4_sharded TR 4
