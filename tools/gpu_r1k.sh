#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export BENCH_TRACE=1 BENCH_HARD_LIMIT_S=120
stat() { python -c "import json,sys
try:
  j=json.loads(sys.stdin.read()); print('value %.1fM e2e %.1fM'%(j['value']/1e6,j['e2e']['value']/1e6), [(k['name'],round(k['ms']*1e3,1),round(k['frac'],3)) for k in j['kernels']], 'large', [(k['name'],round(k['ms']*1e3,1),round(k['frac'],3)) for k in (j.get('mlp_large_batch') or {}).get('kernels',[])])
except Exception as e: print('n/a', e)"; }
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "fused or graph_replay or end_to_end" 2>&1 | tail -n 5
B="python bench.py --cpu-seconds 0 --kernel-reps 5"
for i in 1 2; do timeout 100 $B > gpurun_out/fused_$i.log 2>&1; echo "fused run $i rc=$? $(tail -n 1 gpurun_out/fused_$i.log | stat)"; done
FR_FUSE=0 timeout 100 $B > gpurun_out/unfused.log 2>&1; echo "unfused rc=$? $(tail -n 1 gpurun_out/unfused.log | stat)"
timeout 100 $B --model medium --steps 1000 > gpurun_out/fused_medium.log 2>&1; echo "fused medium rc=$? $(tail -n 1 gpurun_out/fused_medium.log | stat)"
FR_FUSE=0 timeout 100 $B --model medium --steps 1000 > gpurun_out/unfused_medium.log 2>&1; echo "unfused medium rc=$? $(tail -n 1 gpurun_out/unfused_medium.log | stat)"
