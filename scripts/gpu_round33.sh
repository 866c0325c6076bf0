#!/bin/bash
# L2 warm-up in front of the grid walks (LGB_L2_WARM = 1: primitives + camera grid before k_cprimary; 2: + light grids before k_gshadow)
mkdir -p gpurun_out
{
for wm in 0 1 2; do for m in scene_first film_first; do echo "LGB_L2_WARM=$wm:"; LGB_L2_WARM=$wm MODE=$m python scripts/diag_order.py 2>&1 | tail -1; done; done
for wm in 0 1 2; do echo "e2e LGB_L2_WARM=$wm:"; LGB_L2_WARM=$wm python scripts/e2e_breakdown.py mixed4k 2>&1 | tail -2; done
for wm in 0 2; do echo "e2e spheres1m LGB_L2_WARM=$wm:"; LGB_L2_WARM=$wm python scripts/e2e_breakdown.py spheres1m 2>&1 | tail -2; done
} > gpurun_out/r2_v35_l2_warm.txt 2>&1
cat gpurun_out/r2_v35_l2_warm.txt
