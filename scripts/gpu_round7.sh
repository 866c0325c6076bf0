#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
scripts/variants_run.sh mixed4k 2>&1 | tee -a gpurun_out/variants.log
python scripts/profile_kernels.py mixed4k 2>&1 | tee gpurun_out/profile_kernels_${TAG:-r2}.txt
for w in mixed4k spheres1m mesh1m; do LGB_TIMING=1 python scripts/e2e_breakdown.py $w 2>&1 | tail -12; done | tee gpurun_out/e2e_breakdown_${TAG:-r2}.txt
