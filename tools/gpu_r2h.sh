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
for s in 8 10 12 14 16 20; do run n1_s$s 200 python bench.py --gpus 1 --steps 20 --warmup 5 --quick --streams $s; done
run n1_s12_g8 200 python bench.py --gpus 1 --steps 20 --warmup 5 --quick --streams 12 --group 8
echo "=== tests"; timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/t_all_r02h.log 2>&1; echo "exit $?"; tail -n 4 gpurun_out/t_all_r02h.log
