#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "$@" > gpurun_out/$name.log 2>&1; echo "exit $? ($name)"; tail -n ${TAILN:-12} gpurun_out/$name.log; }
run t_layer 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "single_layer or known_answer"
run t_all   900 python -m pytest tests -q -m gpu
