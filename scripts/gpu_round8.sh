#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
for w in mixed4k mesh1m; do LGB_TIMING=1 python scripts/e2e_breakdown.py $w 2>&1 | tail -32; done | tee gpurun_out/e2e_breakdown_${TAG:-r2}.txt
