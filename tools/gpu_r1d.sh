#!/bin/bash
# 8-stream soak, config 5 stress gather (+ncu), config 3 sweep
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "$@" > gpurun_out/$name.log 2>&1; echo "exit $? ($name)"; tail -n ${TAILN:-2} gpurun_out/$name.log | cut -c1-${CUT:-600}; }
B="python bench.py --warmup 50 --cpu-seconds 0 --kernel-reps 5"
run soak8a 120 $B --streams 8 --steps 20000
run soak8b 120 $B --streams 8 --steps 2000
run soak8c 120 $B --streams 8 --steps 2000
CUT=3000 run stress 900 python bench.py --workload stress
CUT=3000 run stress1m 300 python bench.py --workload stress --stress-rows 1000000
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gather_concat -s 6 -c 2 \
  -o gpurun_out/prof_r01b_gather_stress -f python bench.py --workload stress --stress-rows 1000000 --steps 3 > gpurun_out/ncu_gather_stress.log 2>&1
echo "ncu gather exit $?"
CUT=6000 run sweep 900 python bench.py --workload sweep
ls -la gpurun_out/*.ncu-rep
