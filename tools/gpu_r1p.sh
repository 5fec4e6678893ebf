#!/bin/bash
# final refresh of the evidence for the final binary (no ncu: the kernels of the default path are unchanged)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export BENCH_HARD_LIMIT_S=500
run() { name=$1; shift; echo "=== $name"; timeout "$@" > gpurun_out/$name.log 2>&1; echo "exit $? ($name)"; tail -n ${TAILN:-1} gpurun_out/$name.log | cut -c1-${CUT:-300}; }
TAILN=3 run t_all 900 python -m pytest tests -q -m gpu
run smoke 300 python -c "import __graft_entry__ as g; g.smoke()"
run bench_r01 600 python bench.py
run benchref_r01 600 python bench.py --impl reference
run stress_r01 600 python bench.py --workload stress
run bench_medium_r01 300 python bench.py --model medium --cpu-seconds 0 --steps 1000
run shard_gap_r01 300 python tools/shard_gap.py
