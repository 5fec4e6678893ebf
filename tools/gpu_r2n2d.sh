#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export BENCH_HARD_LIMIT_S=300
N=${N:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
run() { name=$1; shift; echo "=== $name"; timeout "$@" > gpurun_out/$name.log 2>&1; echo "exit $? ($name)"; grep '^{' gpurun_out/$name.log | tail -n 1 | python -c "
import sys, json
for l in sys.stdin:
    j = json.loads(l)
    print({k: j.get(k) for k in ('value', 'us_per_batch', 'n_gpus')}, 'e2e', j.get('e2e', {}).get('value'), 'misses', j.get('graph_misses_in_timed_region'))
" || tail -n 5 gpurun_out/$name.log; }
timeout 600 python -m pytest tests/test_round2_gpu.py tests/test_shard.py -q -x -m gpu > gpurun_out/t_shard_n${N}.log 2>&1; echo "tests exit $?"; tail -n 3 gpurun_out/t_shard_n${N}.log
run n${N}_sharded 300 $TR bench.py --gpus $N --steps 20 --warmup 5
[ "$N" = 2 ] && run n${N}_large_b16384 300 $TR bench.py --gpus $N --steps 10 --warmup 3 --model large --batch $((16384 / N)) --rounds 8
[ "$N" = 4 ] && run n${N}_large_b4096 300 $TR bench.py --gpus $N --steps 10 --warmup 3 --model large --batch $((4096 / N)) --rounds 16
[ "$N" = 4 ] && run n${N}_large_b16384 300 $TR bench.py --gpus $N --steps 10 --warmup 3 --model large --batch $((16384 / N)) --rounds 8
true
