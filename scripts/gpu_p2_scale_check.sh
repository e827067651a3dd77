#!/bin/bash
# P2 at scale on one B200: kernels against their host emulation at 1.2 M unknowns, 3D GMRES + Schur against the oracle, and the
# 3D timing twice (iteration counts must repeat).
TAG=${1:-r02s}
mkdir -p gpurun_out
timeout 420 python -m pytest tests/test_gpu_p2.py -m gpu -q -p no:cacheprovider -k "scale or tissue_block" > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc $?" | tee -a gpurun_out/${TAG}_pytest.log
grep -v Warning gpurun_out/${TAG}_pytest.log | tail -n 15
for i in 1 2; do timeout 200 python scripts/perf_p2.py 3d 512 32 > gpurun_out/${TAG}_perf3d_$i.json 2> gpurun_out/${TAG}_perf3d_$i.err; cat gpurun_out/${TAG}_perf3d_$i.json; done
