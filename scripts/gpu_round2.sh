#!/bin/bash
# parity tests, then A/B of the light grids on every config
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
for w in ${WORKLOADS:-mixed4k spheres1m mesh1m cornell simple}; do
  for g in 1 0; do echo "== $w LGB_LIGHT_GRIDS=$g"; LGB_TIMING=1 LGB_LIGHT_GRIDS=$g python scripts/profile_frame.py $w 3 2>&1 | grep -E "frame 2|grids" | cut -c1-150; done
done 2>&1 | tee gpurun_out/grids_ab.log
python scripts/profile_kernels.py mixed4k 2>&1 | tee gpurun_out/profile_kernels_${TAG:-r2}.txt
