#!/bin/bash
# One gpurun call: parity tests in separately time-boxed groups (a hung tcgen05 kernel
# must not eat the rest), smoke, a short bench, and an ncu launch list.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi -L
run() { name=$1; shift; echo "=== $name"; timeout "$@" > gpurun_out/$name.log 2>&1; echo "exit $? ($name)"; tail -n 12 gpurun_out/$name.log; }
run t_gather 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "gather or errors or merge"
run t_fp32   600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "FP32 or (mlp and 1-)"
run t_layer  300 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "single_layer"
run t_all    900 python -m pytest tests -q -m gpu
run smoke    300 python -c "import __graft_entry__ as g; g.smoke()"
run bench    600 python bench.py --steps 500 --warmup 20
run host     300 gpu-fpga-recommendation-system_b200/host/fleetrec_host small 2048 256 4 reference linear tf32
