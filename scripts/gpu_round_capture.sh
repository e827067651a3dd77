#!/bin/bash
# Round capture on one B200: full bench line, ncu launch list of the timed region, ncu --set full of the top kernels.
set -x
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
tail -c 600 gpurun_out/bench_final.json
python bench.py --steps 2 --no-cpu-baseline > gpurun_out/bench_pre.json 2> gpurun_out/bench_pre.err || exit 1
S=$(grep -o '[0-9]* kernel launches' gpurun_out/bench_pre.err | head -1 | cut -d' ' -f1)
echo "launches before timed region: $S"
ncu --metrics gpu__time_duration.sum --clock-control none -s $S -c 1200 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 2 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
python scripts/profile_probe.py 2048 > gpurun_out/plain_probe.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'rows_kernel|spmv_stream_kernel|facet_kernel|multi_dot' -s 3 -c 14 -o gpurun_out/prof_final python scripts/profile_probe.py 2048 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
