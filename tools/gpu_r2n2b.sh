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
run n${N}_waitk 400 $TR bench.py --gpus $N --steps 20 --warmup 5
run n${N}_waitk_s16 400 $TR bench.py --gpus $N --steps 20 --warmup 5 --streams 16
run n${N}_waitk_repl64 400 $TR bench.py --gpus $N --steps 20 --warmup 5 --replicate-mb 64
run n${N}_large_b4096 400 $TR bench.py --gpus $N --steps 10 --warmup 3 --model large --batch $((4096 / N)) --rounds 16
run n${N}_large_b16384 400 $TR bench.py --gpus $N --steps 10 --warmup 3 --model large --batch $((16384 / N)) --rounds 8
run n1_large_quick 400 python bench.py --gpus 1 --steps 10 --warmup 3 --model large --batch 2048 --rounds 8 --quick
( export FLEETREC_LIB=$PWD/gpu-fpga-recommendation-system_b200/libfleetrec_exp.so; FR_SHARD_FOLD=1 run n${N}_fold1 400 $TR bench.py --gpus $N --steps 20 --warmup 5 )
python tools/pcie_probe.py 2>&1 | tail -12
