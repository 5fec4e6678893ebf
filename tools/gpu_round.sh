#!/bin/bash
# One gpurun call: GPU parity tests, smoke, bench (both arms), host driver, then the ncu evidence.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi -L
nproc
run() { name=$1; shift; echo "=== $name"; timeout "$@" > gpurun_out/$name.log 2>&1; echo "exit $? ($name)"; tail -n ${TAILN:-6} gpurun_out/$name.log | cut -c1-2000; }
run t_all    900 python -m pytest tests -q -m gpu
run smoke    300 python -c "import __graft_entry__ as g; g.smoke()"
run bench    600 python bench.py
run benchref 600 python bench.py --impl reference
run host     300 gpu-fpga-recommendation-system_b200/host/fleetrec_host small 2048 256 4 reference linear tf32
bash tools/gpu_profile.sh
