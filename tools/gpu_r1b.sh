#!/bin/bash
# Session-3 GPU call: persistent tcgen05 kernels -- parity first, then per-kernel times, library baseline, bench.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi -L | head -2
run() { name=$1; shift; echo "=== $name"; timeout "$@" > gpurun_out/$name.log 2>&1; echo "exit $? ($name)"; tail -n ${TAILN:-6} gpurun_out/$name.log | cut -c1-1500; }
TAILN=15 run t_tc 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "persistent or single_layer or known_answer or graph_replay"
run t_all 900 python -m pytest tests -q -m gpu
TAILN=12 run sweep 300 python tools/sweep_tiles.py small
run cublas 300 python tools/cublas_ref.py small
run bench 600 python bench.py --cpu-seconds 5
FR_PDL=0 run bench_nopdl 300 python bench.py --cpu-seconds 0 --kernel-reps 5
run bench_s8 300 python bench.py --cpu-seconds 0 --kernel-reps 5 --streams 8
run bench_s2 300 python bench.py --cpu-seconds 0 --kernel-reps 5 --streams 2
