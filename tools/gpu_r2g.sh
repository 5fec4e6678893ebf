#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "$@" > gpurun_out/$name.log 2>&1; echo "exit $? ($name)"; tail -n ${TAILN:-1} gpurun_out/$name.log | cut -c1-${CUT:-600}; }
export FLEETREC_LIB=$PWD/gpu-fpga-recommendation-system_b200/libfleetrec_exp.so
for B in 2048 16384; do
python tools/prof_kernels.py small $B 30
FR_TC_MCAST=1 python tools/prof_kernels.py small $B 30
FR_TC_TILES=256,256,256,2 FR_TC_MCAST=1 python tools/prof_kernels.py small $B 30
done
TAILN=12 run t_exp_r02g 900 python -m pytest tests/test_gpu_parity.py -q -x -m "gpu and experimental"
TAILN=12 run t_shard_r02g 900 env -u FLEETREC_LIB python -m pytest tests/test_round2_gpu.py tests/test_shard.py -q -x -m gpu
