#!/bin/bash
# End-of-round check on a 2-GPU box (gpurun --gpus 2): the whole GPU suite (incl. the 2-GPU tests and the P2 tests), a short
# 1-GPU bench and a 2-GPU bench with the parity check.  Text outputs only.   TAG=$1
TAG=${1:-r02g}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/${TAG}_gpus.log 2>&1
timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest rc $?" | tee -a gpurun_out/${TAG}_pytest.log
tail -5 gpurun_out/${TAG}_pytest.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench1.json 2> gpurun_out/${TAG}_bench1.err
echo "bench1 rc $?"; tail -c 400 gpurun_out/${TAG}_bench1.json
if [ "$(nvidia-smi -L | wc -l)" -ge 2 ]; then
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 \
    bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench2.json 2> gpurun_out/${TAG}_bench2.err
  echo "bench2 rc $?"; tail -c 600 gpurun_out/${TAG}_bench2.json
fi
