#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
for w in ${WORKLOADS:-mixed4k}; do scripts/variants_run.sh $w 2>&1 | tee -a gpurun_out/variants.log; done
python scripts/profile_kernels.py mixed4k 2>&1 | tee gpurun_out/profile_kernels_${TAG:-r2}.txt
for w in spheres1m mesh1m cornell simple; do python scripts/profile_kernels.py $w 2>&1 | tail -8; done | tee gpurun_out/profile_kernels_others_${TAG:-r2}.txt
