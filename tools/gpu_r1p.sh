#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export BENCH_HARD_LIMIT_S=500
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -n 2
timeout 600 python bench.py > gpurun_out/bench_r01.log 2>&1; echo "bench exit $?"; tail -n 1 gpurun_out/bench_r01.log | python -c "
import json,sys
j=json.loads(sys.stdin.read()); print('value %.1f e2e %.1f'%(j['value']/1e6,j['e2e']['value']/1e6), [(k['name'][:10], round(k['ms']*1e3,1)) for k in j['kernels']])"
