#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "graph_replay or end_to_end" 2>&1 | tail -5
for tiles in 128,128,256,1 256,256,256,2 256,256,256,1 128,128,256,2; do
 for streams in 2 4 8; do
  timeout 120 python bench.py --steps 2000 --warmup 30 --cpu-seconds 0 --kernel-reps 5 --tiles $tiles --streams $streams 2>&1 | tail -1 | python -c "
import json,sys
j=json.loads(sys.stdin.read())
print('tiles $tiles streams $streams value %.1f M/s  e2e %.1f M/s  us/step %.2f launches %d'%(j['value']/1e6, j['e2e']['value']/1e6, j['ms_per_step']*1e3, j['gpu_launches']))"
 done
done
FR_GRAPHS=0 timeout 120 python bench.py --steps 2000 --warmup 30 --cpu-seconds 0 --kernel-reps 5 --tiles 128,128,256,1 --streams 4 2>&1 | tail -1 | cut -c1-200
