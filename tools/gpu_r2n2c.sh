#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export BENCH_HARD_LIMIT_S=400
N=${N:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
run() { name=$1; shift; echo "=== $name"; timeout "$@" > gpurun_out/$name.log 2>&1; echo "exit $? ($name)"; grep '^{' gpurun_out/$name.log | tail -n 1 | python -c "
import sys, json
for l in sys.stdin:
    j = json.loads(l)
    print({k: j.get(k) for k in ('value', 'us_per_batch', 'n_gpus')}, 'e2e', j.get('e2e', {}).get('value'), 'misses', j.get('graph_misses_in_timed_region'))
" || tail -n 5 gpurun_out/$name.log; }
TAILN=12 run t_shard_n$N 900 python -m pytest tests/test_round2_gpu.py tests/test_shard.py -q -x -m gpu
tail -n 3 gpurun_out/t_shard_n$N.log
run n${N}_sharded 400 $TR bench.py --gpus $N --steps 20 --warmup 5
run n${N}_sharded_contig 400 $TR bench.py --gpus $N --steps 20 --warmup 5 --plan contiguous
run n${N}_sharded_s16 400 $TR bench.py --gpus $N --steps 20 --warmup 5 --streams 16
run n${N}_sharded_repl64 400 $TR bench.py --gpus $N --steps 20 --warmup 5 --replicate-mb 64
run n${N}_large_b4096 400 $TR bench.py --gpus $N --steps 10 --warmup 3 --model large --batch $((4096 / N)) --rounds 16
