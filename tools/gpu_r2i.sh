#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export BENCH_HARD_LIMIT_S=300
run() { name=$1; shift; echo "=== $name"; timeout "$@" > gpurun_out/$name.log 2>&1; echo "exit $? ($name)"; grep '^{' gpurun_out/$name.log | tail -n 1 | python -c "
import sys, json
for l in sys.stdin:
    j = json.loads(l)
    print({k: j.get(k) for k in ('value', 'us_per_batch')}, 'e2e', j.get('e2e', {}).get('value'))
" || tail -n 5 gpurun_out/$name.log; }
Q="python bench.py --gpus 1 --steps 20 --warmup 5 --quick"
run v_auto 200 $Q
run v_256 200 $Q --tiles 256,256,256,2
run v_512 200 $Q --tiles 512,512,256,2
run v_256_512 200 $Q --tiles 256,512,256,2
export FLEETREC_LIB=$PWD/gpu-fpga-recommendation-system_b200/libfleetrec_exp.so
for kb in 16 24 48 64; do FR_TC_MIN_KB=$kb run v_kb$kb 200 $Q; done
FR_TC_MIN_KB=64 run v_256_kb64 200 $Q --tiles 256,256,256,2
FR_TC_MIN_KB=48 run v_256_kb48 200 $Q --tiles 256,256,256,2
