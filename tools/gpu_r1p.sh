#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 80 python -m pytest tests -q -m gpu -x 2>&1 | tail -n 3
