#!/bin/bash
# P2 row kernel variants on one B200: parity tests with the default, then assembly timings with both register caps.
TAG=${1:-r02q}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_p2.py -m gpu -q -p no:cacheprovider > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc $?" | tee -a gpurun_out/${TAG}_pytest.log
KNP_P2_MINB=6 timeout 200 python -m pytest tests/test_gpu_p2.py -m gpu -q -p no:cacheprovider -k structure > gpurun_out/${TAG}_pytest6.log 2>&1; echo "pytest(minb 6) rc $?" | tee -a gpurun_out/${TAG}_pytest6.log
for mb in 4 6; do
  KNP_P2_MINB=$mb timeout 200 python scripts/perf_p2.py both 512 32 --asm-only > gpurun_out/${TAG}_asm_minb$mb.json 2> gpurun_out/${TAG}_asm_minb$mb.err
  cat gpurun_out/${TAG}_asm_minb$mb.json
done
KNP_P2_MINB=4 timeout 300 python scripts/perf_p2.py both 512 32 > gpurun_out/${TAG}_p2_perf.json 2> gpurun_out/${TAG}_p2_perf.err; cat gpurun_out/${TAG}_p2_perf.json
tail -3 gpurun_out/${TAG}_pytest.log gpurun_out/${TAG}_pytest6.log
