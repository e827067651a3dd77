#!/bin/bash
# Occupancy sweep of the edge-lane row kernel on one B200: rebuilds assembly.cu with different launch bounds / CTA sizes
# and times the assembly on C3 (2D) and C4 (3D).  Usage: gpu_rows_sweep.sh TAG "T M2 M3" ...   (threads, min CTAs 2D, 3D)
TAG=$1; shift
mkdir -p gpurun_out
for v in "$@"; do
  set -- $v
  ( cd knp-emi-cgx_b200/csrc && rm -f assembly.o && make EXTRA="-DEDGE_THREADS=$1 -DEDGE_MIN_CTAS=$2 -DEDGE_MIN_CTAS_3D=$3" > /dev/null 2>&1 ) || { echo "build failed for $v"; continue; }
  echo "== threads $1 min CTAs 2D $2 3D $3" | tee -a gpurun_out/${TAG}_sweep.log
  python scripts/perf_asm.py both 2>&1 | tee -a gpurun_out/${TAG}_sweep.log
done
( cd knp-emi-cgx_b200/csrc && rm -f assembly.o && make > /dev/null 2>&1 )
