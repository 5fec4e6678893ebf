#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export BENCH_HARD_LIMIT_S=500
run() { name=$1; shift; echo "=== $name"; timeout "$@" > gpurun_out/$name.log 2>&1; echo "exit $? ($name)"; tail -n ${TAILN:-1} gpurun_out/$name.log | cut -c1-${CUT:-600}; }
TAILN=25 run t_all_r02e 1500 python -m pytest tests -q -m gpu -x
TAILN=12 CUT=200 run sweep_tiles_r02e 600 python tools/sweep_tiles.py small 2048 4096 16384
( export FLEETREC_LIB=$PWD/gpu-fpga-recommendation-system_b200/libfleetrec_exp.so FR_TC_PROF=1; TAILN=40 CUT=300 run tc_prof_r02e 300 python tools/tc_prof.py small 2048 16384 )
TAILN=1 CUT=1500 run bench_r02e 600 python bench.py --gpus 1 --steps 20 --warmup 5
