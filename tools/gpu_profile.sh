#!/bin/bash
# ncu evidence for profiles/: launch list of the bench command + full captures of the top kernels.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
R=${ROUND:-r01}
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 400 --csv \
  --log-file gpurun_out/launches_$R.csv python bench.py --steps 80 --warmup 5 --cpu-seconds 0 --kernel-reps 2 > gpurun_out/ncu_launch_$R.log 2>&1
echo "launch list exit $?"; tail -n 3 gpurun_out/ncu_launch_$R.log | cut -c1-300
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_linear -s 12 -c 6 \
  -o gpurun_out/prof_${R}_mlp -f python bench.py --steps 10 --warmup 3 --cpu-seconds 0 --kernel-reps 2 > gpurun_out/ncu_mlp_$R.log 2>&1
echo "mlp capture exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gather_concat -s 4 -c 2 \
  -o gpurun_out/prof_${R}_gather -f python bench.py --steps 10 --warmup 3 --cpu-seconds 0 --kernel-reps 2 > gpurun_out/ncu_gather_$R.log 2>&1
echo "gather capture exit $?"
# the large-batch stand-alone gather (uniform indices, 16384 items): last gather launches of the run
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gather_concat -s 60 -c 2 \
  -o gpurun_out/prof_${R}_gather_big -f python bench.py --steps 10 --warmup 3 --cpu-seconds 0 --kernel-reps 2 > gpurun_out/ncu_gather_big_$R.log 2>&1
echo "gather big capture exit $?"
ls -la gpurun_out/*.ncu-rep
