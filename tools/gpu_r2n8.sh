#!/bin/bash
# N-GPU evidence run (N = 4 or 8): sharded small model (driver-style), owner plans, replicated, config 4 (large), config 5 (stress)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export BENCH_HARD_LIMIT_S=400
N=${N:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
run() { name=$1; shift; echo "=== $name"; timeout "$@" > gpurun_out/$name.log 2>&1; echo "exit $? ($name)"; grep '^{' gpurun_out/$name.log | tail -n 1 | python -c "
import sys, json
for l in sys.stdin:
    j = json.loads(l)
    print({k: j.get(k) for k in ('value', 'us_per_batch', 'n_gpus', 'unit')}, 'e2e', j.get('e2e', {}).get('value'), 'misses', j.get('graph_misses_in_timed_region'), 'clocks', j.get('clocks'))
" || tail -n 5 gpurun_out/$name.log; }
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -n 8
run n${N}_sharded 400 $TR bench.py --gpus $N --steps 20 --warmup 5
run n${N}_sharded_contig 400 $TR bench.py --gpus $N --steps 20 --warmup 5 --plan contiguous
run n${N}_sharded_contig_s16 400 $TR bench.py --gpus $N --steps 20 --warmup 5 --plan contiguous --streams 16
run n${N}_replicated 400 $TR bench.py --gpus $N --steps 20 --warmup 5 --shard replicated
run n${N}_large_b4096 400 $TR bench.py --gpus $N --steps 10 --warmup 3 --model large --batch $((4096 / N)) --rounds 16
run n${N}_large_b16384 400 $TR bench.py --gpus $N --steps 10 --warmup 3 --model large --batch $((16384 / N)) --rounds 8
run n${N}_large_b16384_contig 400 $TR bench.py --gpus $N --steps 10 --warmup 3 --model large --batch $((16384 / N)) --rounds 8 --plan contiguous
run n${N}_stress 600 $TR bench.py --gpus $N --workload stress
run n${N}_reference 400 $TR bench.py --impl reference --gpus $N --steps 2 --warmup 1
