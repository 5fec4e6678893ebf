#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python tools/merge_bench.py > gpurun_out/merge_r02.log 2>&1; echo "merge exit $?"; tail -n 1 gpurun_out/merge_r02.log | cut -c1-1800
timeout 300 compute-sanitizer --tool memcheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_memcheck_r02.log 2>&1; echo "memcheck exit $?"; tail -n 4 gpurun_out/sanitizer_memcheck_r02.log
timeout 300 compute-sanitizer --tool racecheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_racecheck_r02.log 2>&1; echo "racecheck exit $?"; tail -n 6 gpurun_out/sanitizer_racecheck_r02.log
timeout 300 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_round2_gpu.py -q -x -m gpu -k "grouped or infer_many or check_indices" > gpurun_out/sanitizer_memcheck_tests_r02.log 2>&1; echo "memcheck tests exit $?"; tail -n 4 gpurun_out/sanitizer_memcheck_tests_r02.log
