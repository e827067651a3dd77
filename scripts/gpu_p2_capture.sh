#!/bin/bash
# P2 evidence on one B200: timings of the P2 path, then (only after that run exited 0 without ncu) the launch list with DRAM
# bytes of the P2 kernels and one full-set capture of the row / facet kernels.  Text outputs only.
TAG=${1:-r02p}
mkdir -p gpurun_out
timeout 400 python scripts/perf_p2.py both 512 32 > gpurun_out/${TAG}_p2_perf.json 2> gpurun_out/${TAG}_p2_perf.err || { tail -5 gpurun_out/${TAG}_p2_perf.err; exit 1; }
cat gpurun_out/${TAG}_p2_perf.json
timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:p2_ -c 28 --csv \
  --log-file gpurun_out/${TAG}_p2_launches.csv python scripts/perf_p2.py both 512 32 --asm-only > gpurun_out/${TAG}_p2_ncu1.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:p2_ -s 4 -c 2 -o /tmp/${TAG}_p2 python scripts/perf_p2.py 3d 512 32 --asm-only > gpurun_out/${TAG}_p2_ncu2.log 2>&1
ncu -i /tmp/${TAG}_p2.ncu-rep --page raw --csv > gpurun_out/${TAG}_p2_full_raw.csv 2>/dev/null
ls -la gpurun_out | tail -8
