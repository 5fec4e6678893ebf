#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_round2_gpu.py -q -x -m gpu -k "reduced_precision or infer_many or grouped or gather_bit_exact" > gpurun_out/t_fp8.log 2>&1; echo "exit $?"; tail -n 25 gpurun_out/t_fp8.log | cut -c1-300
