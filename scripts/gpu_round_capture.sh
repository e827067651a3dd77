#!/bin/bash
# Round capture on one B200 (tag = $1, default r02).  Everything that comes back is a small text file: the .ncu-rep files stay
# on the box (gpurun_out/ is limited to 64 MiB).  Every ncu command runs only after the same command exited 0 without ncu.
#   1. full bench line (with the CPU baseline leg)                          -> ${TAG}_bench.json
#   2. ncu launch list of the timed region of the same bench                -> ${TAG}_launches.csv
#   3. ncu --set full of gate / facet / rows / SpMV(A) / the first kernels of the cycle -> ${TAG}_full_raw.csv,
#      source page of the row kernel                                        -> ${TAG}_rows_src.csv
#   4. DRAM bytes + time of every kernel of one assembly + SpMV + preconditioner application -> ${TAG}_traffic.csv
TAG=${1:-r02}
set -x
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err || { tail -5 gpurun_out/${TAG}_bench.err; exit 1; }
tail -c 300 gpurun_out/${TAG}_bench.json
python bench.py --steps 2 --no-cpu-baseline --skip-c4 --skip-parity > gpurun_out/${TAG}_bench_pre.json 2> gpurun_out/${TAG}_bench_pre.err || exit 1
S=$(grep -o '[0-9]* kernel launches' gpurun_out/${TAG}_bench_pre.err | head -1 | cut -d' ' -f1)
echo "launches before timed region: $S"
ncu --metrics gpu__time_duration.sum --clock-control none -s $S -c 1000 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --no-cpu-baseline --skip-c4 --skip-parity > gpurun_out/${TAG}_ncu_launch.log 2>&1
export KNP_PC_GRAPH=0
python scripts/profile_probe.py c3 2048 > gpurun_out/${TAG}_probe.log 2>&1 || { tail -5 gpurun_out/${TAG}_probe.log; exit 1; }
cat gpurun_out/${TAG}_probe.log
S2=$(grep -o 'second repetition [0-9]*' gpurun_out/${TAG}_probe.log | cut -d' ' -f3)
C2=$(grep -o 'per repetition [0-9]*' gpurun_out/${TAG}_probe.log | cut -d' ' -f3)
ncu --set full --clock-control none --import-source on -s $((S2 + 1)) -c 10 -o /tmp/${TAG}_prof python scripts/profile_probe.py c3 2048 > gpurun_out/${TAG}_ncu_full.log 2>&1
ncu -i /tmp/${TAG}_prof.ncu-rep --page raw --csv > gpurun_out/${TAG}_full_raw.csv 2>/dev/null
ncu -i /tmp/${TAG}_prof.ncu-rep --page source --csv --print-source cuda,sass --kernel-name regex:rows_edge > gpurun_out/${TAG}_rows_src.csv 2>/dev/null
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s $((S2 + 1)) -c $C2 --csv --log-file gpurun_out/${TAG}_traffic.csv python scripts/profile_probe.py c3 2048 > gpurun_out/${TAG}_ncu_traffic.log 2>&1
ls -la gpurun_out; du -sh gpurun_out
