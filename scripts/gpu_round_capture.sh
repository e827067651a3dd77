#!/bin/bash
# Round capture on one B200 (tag = $1, default r02): full bench line, ncu launch list of the timed region, ncu --set full of one
# gate + assembly + SpMV + preconditioner application.  Every ncu command runs only after the same command exited 0 without ncu.
TAG=${1:-r02}
set -x
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err || exit 1
tail -c 400 gpurun_out/${TAG}_bench.json
python bench.py --steps 2 --no-cpu-baseline --skip-c4 --skip-parity > gpurun_out/${TAG}_bench_pre.json 2> gpurun_out/${TAG}_bench_pre.err || exit 1
S=$(grep -o '[0-9]* kernel launches' gpurun_out/${TAG}_bench_pre.err | head -1 | cut -d' ' -f1)
echo "launches before timed region: $S"
ncu --metrics gpu__time_duration.sum --clock-control none -s $S -c 1000 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --no-cpu-baseline --skip-c4 --skip-parity > gpurun_out/${TAG}_ncu_launch.log 2>&1
export KNP_PC_GRAPH=0
python scripts/profile_probe.py c3 2048 > gpurun_out/${TAG}_probe.log 2>&1 || exit 1
cat gpurun_out/${TAG}_probe.log
S2=$(grep -o 'second repetition [0-9]*' gpurun_out/${TAG}_probe.log | cut -d' ' -f3)
C2=$(grep -o 'per repetition [0-9]*' gpurun_out/${TAG}_probe.log | cut -d' ' -f3)
ncu --set full --clock-control none --import-source on -s $((S2 + 1)) -c $C2 -o gpurun_out/${TAG}_prof python scripts/profile_probe.py c3 2048 > gpurun_out/${TAG}_ncu_full.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_full.log
ls -la gpurun_out/${TAG}_prof.ncu-rep
