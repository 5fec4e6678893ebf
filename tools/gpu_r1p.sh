#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
R=r01
export BENCH_HARD_LIMIT_S=300
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 480 --csv \
  --log-file gpurun_out/launches_$R.csv python bench.py --steps 100 --warmup 5 --cpu-seconds 0 --kernel-reps 2 --gather-batch 2048 > gpurun_out/ncu_launch_$R.log 2>&1
echo "launch list exit $?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"tc_linear|gather_concat" -s 40 -c 8 \
  -o gpurun_out/prof_${R}_step -f python bench.py --steps 10 --warmup 3 --cpu-seconds 0 --kernel-reps 2 --gather-batch 2048 > gpurun_out/ncu_step_$R.log 2>&1
echo "step capture exit $?"
