#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
for m in scene_first film_first; do MODE=$m python scripts/diag_order.py 2>&1 | tail -1; done
python scripts/profile_kernels.py mixed4k | cut -c1-60
python scripts/profile_kernels.py spheres1m | cut -c1-60
