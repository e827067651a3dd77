#!/bin/bash
TAG=${1:-r02v}
mkdir -p gpurun_out
run() { echo "=== $*" >> gpurun_out/${TAG}_diag.log; env "$@" timeout 120 python scripts/diag_p2_pc.py 32 >> gpurun_out/${TAG}_diag.log 2>&1; }
run DIAG_PRE_ASSEMBLE=7
run DIAG_PRE_ASSEMBLE=1
run DIAG_PRE_ASSEMBLE=7
run DIAG_PRE_ASSEMBLE=3
grep -v Warning gpurun_out/${TAG}_diag.log | grep "===\|^B r:\|iterations" | cut -c1-300
