#!/bin/bash
# Stages the reference's OWN driver scripts and CI configs (unmodified, not committed: baseline/_ref is git-ignored but travels
# with gpurun) so that tests/test_gpu_reference_scripts.py can run them on the GPU box, where /root/reference does not exist.
set -e
R=${1:-/root/reference}
D=$(dirname "$0")/../baseline/_ref
mkdir -p "$D/tests/KNPEMI" "$D/src/CGx/KNPEMI/configs/tests"
cp "$R"/tests/KNPEMI/electric_potential_norms_*_solver.py "$D/tests/KNPEMI/"
cp "$R"/src/CGx/KNPEMI/configs/tests/*.yaml "$D/src/CGx/KNPEMI/configs/tests/"
ls -R "$D" | head -20
