#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
for w in ${WORKLOADS:-mixed4k spheres1m mesh1m cornell simple}; do
  for g in 1 0; do echo "== $w LGB_CAMERA_GRID=$g"; LGB_TIMING=1 LGB_CAMERA_GRID=$g python scripts/profile_frame.py $w 3 2>&1 | grep -E "frame 2|grid" | cut -c1-150; done
done 2>&1 | tee gpurun_out/camgrid_ab.log
for sh in 1 3; do echo "== mixed4k LGB_CAM_SHIFT=$sh"; LGB_TIMING=1 LGB_CAM_SHIFT=$sh python scripts/profile_frame.py mixed4k 3 2>&1 | grep -E "frame 2|grid" | cut -c1-150; done 2>&1 | tee -a gpurun_out/camgrid_ab.log
python scripts/profile_kernels.py mixed4k 2>&1 | tee gpurun_out/profile_kernels_${TAG:-r2}.txt
