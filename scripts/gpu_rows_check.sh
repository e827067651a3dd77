#!/bin/bash
# Row-kernel development loop on one B200: parity tests, assembly timing of both row kernels on C3 / C4, and one
# `ncu --set full` capture of the edge-lane kernel per dimension (only after the same command exited 0 without ncu).
TAG=${1:-rows}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc $?"; tail -5 gpurun_out/${TAG}_pytest.log
python scripts/perf_asm.py both > gpurun_out/${TAG}_asm_edge.log 2>&1 || { tail -5 gpurun_out/${TAG}_asm_edge.log; exit 1; }
cat gpurun_out/${TAG}_asm_edge.log
KNP_ROWS=scan python scripts/perf_asm.py both > gpurun_out/${TAG}_asm_scan.log 2>&1; cat gpurun_out/${TAG}_asm_scan.log
for d in 2d 3d; do
  ncu --set full --clock-control none --import-source on -k regex:rows_edge -s 3 -c 1 -o gpurun_out/${TAG}_prof_$d python scripts/perf_asm.py $d > gpurun_out/${TAG}_ncu_$d.log 2>&1
  ncu -i gpurun_out/${TAG}_prof_$d.ncu-rep --page raw --csv > gpurun_out/${TAG}_raw_$d.csv 2>/dev/null
  ncu -i gpurun_out/${TAG}_prof_$d.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/${TAG}_src_$d.csv 2>/dev/null
done
ls -la gpurun_out | tail -12
du -sh gpurun_out
