#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "$@" > gpurun_out/$name.log 2>&1; echo "exit $? ($name)"; tail -n ${TAILN:-1} gpurun_out/$name.log | cut -c1-${CUT:-600}; }
export FLEETREC_LIB=$PWD/gpu-fpga-recommendation-system_b200/libfleetrec_exp.so FR_TC_PROF=1 FR_TC_NOSTORE=1
TAILN=40 CUT=300 run tc_prof_r02f 300 python tools/tc_prof.py small 16384
FR_TC_TILES=256,256,256,2 TAILN=40 CUT=300 run tc_prof_r02f_256 300 python tools/tc_prof.py small 16384
