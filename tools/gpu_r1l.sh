#!/bin/bash
# is the step rate bounded by kernel dispatch?  same chain at small batches, 12 workers
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export BENCH_HARD_LIMIT_S=120
stat() { python -c "import json,sys
try:
  j=json.loads(sys.stdin.read()); print('value %.1fM us/step %.2f host_us %s'%(j['value']/1e6,j['ms_per_step']*1e3,j['host_enqueue_us_per_step']), [(k['name'],round(k['ms']*1e3,1)) for k in j['kernels']])
except Exception as e: print('n/a', e)"; }
for b in 128 512 1024 2048 4096; do
  python bench.py --cpu-seconds 0 --kernel-reps 3 --batch $b --gather-batch 4096 --steps 4000 > gpurun_out/b$b.log 2>&1; echo "batch $b: $(tail -n 1 gpurun_out/b$b.log | stat)"
done
FR_FUSE=1 python bench.py --cpu-seconds 0 --kernel-reps 3 --batch 128 --gather-batch 4096 --steps 4000 > gpurun_out/b128f.log 2>&1; echo "batch 128 fused: $(tail -n 1 gpurun_out/b128f.log | stat)"
