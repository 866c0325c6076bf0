#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
for v in 1 0; do echo "== LGB_SETUP_IN_PRIMARY=$v"; for w in mixed4k spheres1m; do LGB_SETUP_IN_PRIMARY=$v python scripts/profile_frame.py $w 4 | tail -1 | cut -c1-60; done; done
LGB_SETUP_IN_PRIMARY=1 python scripts/profile_kernels.py mixed4k 2>&1 | tail -6
