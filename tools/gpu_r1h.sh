#!/bin/bash
# persistent-grid cap sweep (fewer, longer-lived clusters) at 8 worker streams
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
B="python bench.py --warmup 50 --cpu-seconds 0 --kernel-reps 5 --steps 2000"
for cap in 0 4 8 12 16 24 32; do
  FR_TC_MAX_CLUSTERS=$cap timeout 60 $B > gpurun_out/cap$cap.log 2>&1
  echo "cap=$cap rc=$? $(tail -n 1 gpurun_out/cap$cap.log | python -c "import json,sys
try:
  j=json.loads(sys.stdin.read()); print('value %.1fM e2e %.1fM'%(j['value']/1e6,j['e2e']['value']/1e6), [(k['name'],round(k['ms']*1e3,1)) for k in j['kernels']])
except Exception as e: print('n/a',e)")"
done
for s in 12 16; do
  timeout 60 $B --streams $s > gpurun_out/s$s.log 2>&1
  echo "streams=$s rc=$? $(tail -n 1 gpurun_out/s$s.log | python -c "import json,sys
try:
  j=json.loads(sys.stdin.read()); print('value %.1fM e2e %.1fM'%(j['value']/1e6,j['e2e']['value']/1e6))
except Exception as e: print('n/a',e)")"
done
FR_TC_MAX_CLUSTERS=16 timeout 60 $B --streams 12 > gpurun_out/cap16s12.log 2>&1; echo "cap16 s12 $(tail -n 1 gpurun_out/cap16s12.log | cut -c1-120)"
