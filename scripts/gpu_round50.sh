#!/bin/bash
mkdir -p gpurun_out
{
echo "== base"; python scripts/profile_kernels.py mixed4k | cut -c1-60
echo "== LGB_SETUP_IN_PRIMARY=0"; LGB_SETUP_IN_PRIMARY=0 python scripts/profile_kernels.py mixed4k | cut -c1-60
for sh in 1 3; do echo "== LGB_CAM_SHIFT=$sh"; LGB_CAM_SHIFT=$sh python scripts/profile_kernels.py mixed4k | cut -c1-60; done
for r in 512 2048; do echo "== LGB_GRID_RES=$r"; LGB_GRID_RES=$r python scripts/profile_kernels.py mixed4k | cut -c1-60; done
for c in 0.5 2.0; do echo "== LGB_GRID_CELL=$c"; LGB_GRID_CELL=$c python scripts/profile_kernels.py mixed4k | cut -c1-60; done
} > gpurun_out/r2_v52_runtime_sweeps.txt 2>&1
cat gpurun_out/r2_v52_runtime_sweeps.txt
