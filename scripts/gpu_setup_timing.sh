#!/bin/bash
# phases of the preconditioner setup on C3 (KNP_AMG_TIMING=1), device and host hierarchy setup
mkdir -p gpurun_out
for where in device host; do
  KNP_AMG_SETUP=$where KNP_AMG_TIMING=1 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --skip-c4 --skip-parity \
    > gpurun_out/setup_$where.json 2> gpurun_out/setup_$where.err
  grep -E "pc setup|amg|setup of" gpurun_out/setup_$where.err > gpurun_out/setup_$where.log
done
tail -n 60 gpurun_out/setup_device.log
