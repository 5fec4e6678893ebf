#!/bin/bash
# 8-stream hang triage + ncu full captures of the persistent kernels
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "$@" > gpurun_out/$name.log 2>&1; echo "exit $? ($name)"; tail -n ${TAILN:-2} gpurun_out/$name.log | cut -c1-${CUT:-400}; }
B="python bench.py --steps 1000 --warmup 20 --cpu-seconds 0 --kernel-reps 3"
FR_PDL=0 run s8_nopdl 90 $B --streams 8
CUDA_DEVICE_MAX_CONNECTIONS=32 run s8_conn32 90 $B --streams 8
run s6 90 $B --streams 6
FR_GRAPHS=0 run s8_nograph 90 $B --streams 8
run s8 90 $B --streams 8
run t128 90 $B --tiles 128,128,256,2
run t256_128 90 $B --tiles 256,128,256,2
nvidia-smi --query-gpu=utilization.gpu,memory.used --format=csv,noheader
for bsz in 16384 2048; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_linear -c 9 \
    -o gpurun_out/prof_r01b_mlp_B$bsz -f python tools/prof_kernels.py small $bsz 1 > gpurun_out/ncu_mlp_B$bsz.log 2>&1
  echo "ncu mlp B=$bsz exit $?"; tail -n 2 gpurun_out/ncu_mlp_B$bsz.log | cut -c1-300
done
ls -la gpurun_out/*.ncu-rep
