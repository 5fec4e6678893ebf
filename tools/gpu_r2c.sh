#!/bin/bash
# Round 2, run C: GPU tests, TMA probe (store / box-size cost), per-kernel times, driver-style bench.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export BENCH_HARD_LIMIT_S=500
run() { name=$1; shift; echo "=== $name"; timeout "$@" > gpurun_out/$name.log 2>&1; echo "exit $? ($name)"; tail -n ${TAILN:-1} gpurun_out/$name.log | cut -c1-${CUT:-600}; }
TAILN=15 run t_all_r02c 1500 python -m pytest tests -q -m gpu -x
TAILN=40 CUT=200 run tma_probe_r02c 300 tools/tma_probe
TAILN=40 CUT=200 run sweep_tiles_r02c 600 python tools/sweep_tiles.py small 2048 4096 16384
TAILN=1 CUT=1500 run bench_r02c 600 python bench.py --gpus 1 --steps 20 --warmup 5
