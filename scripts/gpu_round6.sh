#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
for w in ${WORKLOADS:-mixed4k}; do scripts/variants_run.sh $w 2>&1 | tee -a gpurun_out/variants.log; done
python scripts/big_frame.py 2>&1 | tee gpurun_out/big_frame_${TAG:-r2}.txt
