#!/bin/bash
# The round's evidence run on one B200: all GPU tests, smoke, both bench arms, the other configs,
# then the ncu launch list + full captures.  ROUND=r01 bash tools/gpu_final.sh
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
R=${ROUND:-r01}
export BENCH_HARD_LIMIT_S=500
run() { name=$1; shift; echo "=== $name"; timeout "$@" > gpurun_out/$name.log 2>&1; echo "exit $? ($name)"; tail -n ${TAILN:-1} gpurun_out/$name.log | cut -c1-${CUT:-400}; }
TAILN=3 run t_all 900 python -m pytest tests -q -m gpu
run smoke 300 python -c "import __graft_entry__ as g; g.smoke()"
run bench_$R 600 python bench.py
run benchref_$R 600 python bench.py --impl reference
run host 300 gpu-fpga-recommendation-system_b200/host/fleetrec_host small 2048 256 4 reference linear tf32
run stress_$R 600 python bench.py --workload stress
run sweep_$R 600 python bench.py --workload sweep
run cublas_$R 300 python tools/cublas_ref.py small
FR_CHAIN=1 run bench_chain_$R 300 python bench.py --cpu-seconds 0
run bench_medium_$R 300 python bench.py --model medium --cpu-seconds 0 --steps 1000
FR_CHAIN=1 FR_CHAIN_PROF=1 run chain_timeline_$R 120 python tools/chain_timeline.py small 2048
run pcie_$R 200 python tools/pcie_probe.py
# ncu: launch list of the bench command, then full captures
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 480 --csv \
  --log-file gpurun_out/launches_$R.csv python bench.py --steps 100 --warmup 5 --cpu-seconds 0 --kernel-reps 2 > gpurun_out/ncu_launch_$R.log 2>&1
echo "launch list exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"tc_linear|gather_concat" -s 40 -c 8 \
  -o gpurun_out/prof_${R}_step -f python bench.py --steps 10 --warmup 3 --cpu-seconds 0 --kernel-reps 2 > gpurun_out/ncu_step_$R.log 2>&1
echo "step capture exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_linear -c 6 \
  -o gpurun_out/prof_${R}_mlp_B16384 -f python tools/prof_kernels.py small 16384 1 > gpurun_out/ncu_mlp16k_$R.log 2>&1
echo "mlp 16384 capture exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gather_concat -s 6 -c 2 \
  -o gpurun_out/prof_${R}_gather_stress -f python bench.py --workload stress --stress-rows 1000000 --steps 3 > gpurun_out/ncu_gstress_$R.log 2>&1
echo "gather stress capture exit $?"
ls -la gpurun_out/*.ncu-rep
