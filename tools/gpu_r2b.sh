#!/bin/bash
# Round 2, run B: all GPU tests (incl. the experiments build in a child process), TMA feed probe, tile sweep, cuBLAS reference.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export BENCH_HARD_LIMIT_S=500
run() { name=$1; shift; echo "=== $name"; timeout "$@" > gpurun_out/$name.log 2>&1; echo "exit $? ($name)"; tail -n ${TAILN:-1} gpurun_out/$name.log | cut -c1-${CUT:-600}; }
TAILN=15 run t_all_r02b 1200 python -m pytest tests -q -m gpu -x
TAILN=30 CUT=200 run tma_probe_r02 300 tools/tma_probe
TAILN=40 CUT=200 run sweep_tiles_r02 600 python tools/sweep_tiles.py small 2048 4096 16384
TAILN=2 CUT=3000 run cublas_r02 300 python tools/cublas_ref.py small
