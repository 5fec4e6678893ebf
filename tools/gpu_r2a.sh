#!/bin/bash
# Round 2, run A: all GPU tests, smoke, the driver's bench command.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export BENCH_HARD_LIMIT_S=500
run() { name=$1; shift; echo "=== $name"; timeout "$@" > gpurun_out/$name.log 2>&1; echo "exit $? ($name)"; tail -n ${TAILN:-1} gpurun_out/$name.log | cut -c1-${CUT:-600}; }
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv
TAILN=15 run t_all_r02a 900 python -m pytest tests -q -m gpu -x
run smoke_r02a 300 python -c "import __graft_entry__ as g; g.smoke()"
BENCH_TRACE=1 TAILN=3 CUT=3000 run bench_r02a 600 python bench.py --gpus 1 --steps 20 --warmup 5
