#!/bin/bash
# one GPU session: parity tests, per-frame timing of the in-tree library and of every build/lib_*.so variant, launch list
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
for w in ${WORKLOADS:-mixed4k}; do scripts/variants_run.sh $w 2>&1 | tee -a gpurun_out/variants.log; done
scripts/launch_times.sh mixed4k ${TAG:-r2} 2>&1 | tee gpurun_out/launch_summary_${TAG:-r2}.txt
