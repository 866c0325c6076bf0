#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
for w in mixed4k spheres1m mesh1m; do LGB_TIMING=1 python scripts/e2e_breakdown.py $w 2>&1 | grep -E " e2e |ensure_bvh|light grids\] [0-9]|lgb_scene_create\]|camera grid" | tail -7; done | tee gpurun_out/e2e_breakdown_${TAG:-r2}.txt
python scripts/profile_kernels.py mixed4k 2>&1 | tee gpurun_out/profile_kernels_${TAG:-r2}.txt
