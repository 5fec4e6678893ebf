#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 300 python tools/shard_gap.py 2>&1 | tail -n 8
echo "--- no flag kernel"; FR_SHARD_NOWAIT=2 timeout 300 python tools/shard_gap.py 2>&1 | tail -n 4
