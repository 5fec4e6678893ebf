#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 200 python tools/pcie_probe.py 2>&1 | tail -n 24
