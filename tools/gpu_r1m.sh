#!/bin/bash
# 512-wide pair tiles: parity, then tile-configuration sweep at 12 workers
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export BENCH_HARD_LIMIT_S=120
stat() { python -c "import json,sys
try:
  j=json.loads(sys.stdin.read()); print('value %.1fM e2e %.1fM'%(j['value']/1e6,j['e2e']['value']/1e6), [(k['name'][:10],round(k['ms']*1e3,1)) for k in j['kernels']], 'B16384', [(round(k['ms']*1e3,1),round(k['frac'],2)) for k in (j.get('mlp_large_batch') or {}).get('kernels',[])])
except Exception as e: print('n/a', e)"; }
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "persistent or single_layer or known_answer" 2>&1 | tail -n 4
for t in 256,256,256,2 512,512,256,2 512,256,256,2 256,512,256,2; do
  python bench.py --cpu-seconds 0 --kernel-reps 5 --tiles $t > gpurun_out/tiles_$t.log 2>&1; echo "tiles $t: $(tail -n 1 gpurun_out/tiles_$t.log | stat)"
done
python bench.py --cpu-seconds 0 --kernel-reps 5 --tiles 512,512,256,2 --model medium --steps 1000 > gpurun_out/tiles_medium512.log 2>&1; echo "medium 512: $(tail -n 1 gpurun_out/tiles_medium512.log | stat)"
python bench.py --cpu-seconds 0 --kernel-reps 5 --model medium --steps 1000 > gpurun_out/tiles_medium256.log 2>&1; echo "medium 256: $(tail -n 1 gpurun_out/tiles_medium256.log | stat)"
