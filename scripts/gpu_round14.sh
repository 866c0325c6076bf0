#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
for w in mixed4k spheres1m; do scripts/variants_run.sh $w; done
python scripts/profile_kernels.py mixed4k 2>&1 | tail -6
