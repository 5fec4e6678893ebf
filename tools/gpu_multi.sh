#!/bin/bash
# The round's N-GPU evidence (N = 2, 4 or 8): N=8 bash tools/gpu_multi.sh   (under `gpurun --gpus N`)
#   - N = 2: the sharded GPU tests on real peers
#   - the driver's command under torchrun (small model, table-sharded, packed index rows) and with int32 rows
#   - config 4: the large model sharded, B_global 4096 and 16384
#   - N = 8: config 5 (stress lookup), and the reference arm
# Every run has its own hard limit (bench.py leaves after BENCH_HARD_LIMIT_S without a result): a hung collective costs
# N x the box time.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export BENCH_HARD_LIMIT_S=${BENCH_HARD_LIMIT_S:-90}
N=${N:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
run() { name=$1; shift; echo "=== $name"; timeout "$@" > gpurun_out/$name.log 2>&1; echo "exit $? ($name)"; grep '^{' gpurun_out/$name.log | tail -n 1 | python -c "
import sys, json
for l in sys.stdin:
    j = json.loads(l)
    print({k: j.get(k) for k in ('value', 'us_per_batch', 'n_gpus', 'unit')}, 'e2e', j.get('e2e', {}).get('value'), 'misses', j.get('graph_misses_in_timed_region'), 'clocks', j.get('clocks'))
" || tail -n 5 gpurun_out/$name.log; }
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -n 8
if [ "$N" = 2 ]; then
  timeout 600 python -m pytest tests/test_round2_gpu.py tests/test_shard.py -q -x -m gpu > gpurun_out/t_shard_n${N}.log 2>&1
  echo "tests exit $?"; tail -n 3 gpurun_out/t_shard_n${N}.log
fi
run n${N}_packed 120 $TR bench.py --gpus $N --steps 20 --warmup 5
run n${N}_sharded 120 $TR bench.py --gpus $N --steps 20 --warmup 5 --index-format i32
run n${N}_large_b4096 200 $TR bench.py --gpus $N --steps 10 --warmup 3 --model large --batch $((4096 / N)) --rounds 16
run n${N}_large_b16384 200 $TR bench.py --gpus $N --steps 10 --warmup 3 --model large --batch $((16384 / N)) --rounds 8
if [ "$N" = 8 ]; then
  BENCH_HARD_LIMIT_S=400 run n${N}_stress 420 $TR bench.py --gpus $N --workload stress
  BENCH_HARD_LIMIT_S=300 run n${N}_reference 320 $TR bench.py --impl reference --gpus $N --steps 2 --warmup 1
fi
true
