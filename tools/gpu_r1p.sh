#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export BENCH_HARD_LIMIT_S=300
timeout 600 python -m pytest tests -q -m gpu -x -k "gather or end_to_end or reduced_precision or merge or shard or smoke" 2>&1 | tail -n 4
summ='import json,sys
j=json.loads(sys.stdin.read()); print("value %.1f e2e %.1f M/s  step %.2f us" % (j["value"]/1e6, j["e2e"]["value"]/1e6, j["ms_per_step"]*1e3), [(k["name"][:12], round(k["ms"]*1e3,2)) for k in j["kernels"]], j["gather_standalone"]["ms"]*1e3, j["gather_standalone"]["frac"])'
timeout 400 python bench.py --cpu-seconds 0 > gpurun_out/bench_tmp.log 2>&1; echo "exit $?"; tail -n 1 gpurun_out/bench_tmp.log | python -c "$summ"
timeout 400 python bench.py --workload stress --steps 20 > gpurun_out/bench_tmp2.log 2>&1; echo "exit $?"; tail -n 1 gpurun_out/bench_tmp2.log | python -c "import json,sys; j=json.loads(sys.stdin.read()); print(j['uniform'], j['zipf'])"
