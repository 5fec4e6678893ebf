#!/bin/bash
# The round's 1-GPU evidence: all GPU tests, smoke, both bench arms as the driver runs them, the config-3 sweep, the C++
# host, cuBLAS on the same shapes, then the ncu launch list and full captures.  ROUND=r02 bash tools/gpu_evidence.sh
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
R=${ROUND:-r02}
export BENCH_HARD_LIMIT_S=500
run() { name=$1; shift; echo "=== $name"; timeout "$@" > gpurun_out/$name.log 2>&1; echo "exit $? ($name)"; tail -n ${TAILN:-1} gpurun_out/$name.log | cut -c1-${CUT:-300}; }
TAILN=3 run t_all_$R 1500 python -m pytest tests -q -m gpu
run smoke_$R 300 python -c "import __graft_entry__ as g; g.smoke()"
run bench_$R 600 python bench.py --gpus 1 --steps 20 --warmup 5
run benchref_$R 600 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1
run sweep_$R 600 python bench.py --workload sweep
run host_$R 300 gpu-fpga-recommendation-system_b200/host/fleetrec_host small 2048 256 4 reference linear tf32
run cublas_$R 300 python tools/cublas_ref.py small
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 480 --csv \
  --log-file gpurun_out/launches_$R.csv python bench.py --steps 2 --warmup 1 --quick > gpurun_out/ncu_launch_$R.log 2>&1
echo "launch list exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"tc_pair|gather_concat" -s 40 -c 8 \
  -o gpurun_out/prof_${R}_step -f python bench.py --steps 1 --warmup 1 --quick > gpurun_out/ncu_step_$R.log 2>&1
echo "step capture exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_pair -c 6 \
  -o gpurun_out/prof_${R}_mlp_B16384 -f python tools/prof_kernels.py small 16384 1 > gpurun_out/ncu_mlp16k_$R.log 2>&1
echo "mlp 16384 capture exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gather_concat -s 6 -c 2 \
  -o gpurun_out/prof_${R}_gather_stress -f python bench.py --workload stress --stress-rows 1000000 --stress-steps 3 > gpurun_out/ncu_gstress_$R.log 2>&1
echo "gather stress capture exit $?"
ls -la gpurun_out/*_$R*.ncu-rep
