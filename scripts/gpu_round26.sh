#!/bin/bash
# warp-per-cell bitonic sort of the grid cells + sub-trees of <= 256 items finished by one warp: parity, builder timing, e2e of C2 / C4 / C5
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
{
LGB_TIMING=1 python scripts/e2e_breakdown.py mesh1m
LGB_TIMING=1 python scripts/e2e_breakdown.py spheres1m
LGB_TIMING=1 python scripts/e2e_breakdown.py mixed4k
} > gpurun_out/r2_v30_e2e_breakdown.txt 2>&1
grep -v "light [0-9]" gpurun_out/r2_v30_e2e_breakdown.txt | grep -v "it0" | tail -50
grep "light [0-9]" gpurun_out/r2_v30_e2e_breakdown.txt | tail -12
