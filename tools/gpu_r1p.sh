#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 240 compute-sanitizer --tool racecheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_racecheck_r01.log 2>&1; echo "racecheck exit $?"; grep -v "^=========     at\|^=========         in" gpurun_out/sanitizer_racecheck_r01.log | cut -c1-230 | tail -n 12
