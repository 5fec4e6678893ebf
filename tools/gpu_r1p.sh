#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export BENCH_HARD_LIMIT_S=300
timeout 400 python bench.py --cpu-seconds 0 > gpurun_out/bench_tmp.log 2>&1; echo "exit $?"; tail -n 1 gpurun_out/bench_tmp.log | python -c "
import json,sys
j=json.loads(sys.stdin.read()); print('value %.1f e2e %.1f'%(j['value']/1e6,j['e2e']['value']/1e6)); 
for k in j['kernels']: print(k['name'], round(k['ms']*1e3,1), k.get('sms_occupied'), round(k.get('frac_of_occupied_sms',0),3), {a:round(b,4) for a,b in k.get('at_step_occupancy',{}).items() if a in ('achieved','frac','ms_per_launch_effective')})"
tail -n 5 gpurun_out/bench_tmp.log | grep -v "^{" | cut -c1-300
