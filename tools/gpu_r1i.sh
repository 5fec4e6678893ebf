#!/bin/bash
# soak of the default configuration (no programmatic edges, ganged short tiles) + PDL hang diagnosis with the watchdog build
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export BENCH_TRACE=1
B="python bench.py --warmup 50 --cpu-seconds 0 --kernel-reps 3 --steps 3000 --gather-batch 2048"
stat() { python -c "import json,sys
try:
  j=json.loads(sys.stdin.read()); print('value %.1fM e2e %.1fM'%(j['value']/1e6,j['e2e']['value']/1e6))
except Exception as e: print('n/a')"; }
soak() { name=$1; n=$2; shift 2; ok=0; bad=0; for i in $(seq $n); do timeout 40 "$@" > gpurun_out/${name}_$i.log 2>&1; rc=$?; if [ $rc -eq 0 ]; then ok=$((ok+1)); else bad=$((bad+1)); echo "  $name run $i rc=$rc: $(grep -E 'watchdog|rror' gpurun_out/${name}_$i.log | tail -n 2 | cut -c1-400)"; fi; done; echo "$name ok=$ok bad=$bad $(tail -n 1 gpurun_out/${name}_1.log | stat)"; }
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "persistent or single_layer or known_answer or graph_replay or end_to_end" 2>&1 | tail -n 2
soak def_s8 4 $B --streams 8
soak def_s12 3 $B --streams 12
soak def_s16 2 $B --streams 16
soak def_s4 2 $B --streams 4
FR_TC_MIN_KB=1 soak nogang_s8 1 $B --streams 8
FR_PDL=6 soak pdl6_s12 4 $B --streams 12
FR_PDL=7 soak pdl7_s8 3 $B --streams 8
FR_PDL=7 FR_GRAPHS=0 soak pdl7_nograph_s8 3 $B --streams 8
