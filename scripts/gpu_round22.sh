#!/bin/bash
# sphere early-out (parity), and: does the per-thread stack size the driver lays local memory out with explain the two states of k_cprimary?
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
{
for lim in "" 256 768 1024 1280 2048 4096; do
  echo "== LGB_STACK_LIMIT=${lim:-unset}"
  LGB_STACK_LIMIT=$lim python scripts/profile_kernels.py mixed4k 2>&1 | grep -E "lgb_init|k_cprimary|k_gshadow" | cut -c1-60
  LGB_STACK_LIMIT=$lim DIAG_STEPS=plain python scripts/diag_bench_gap.py 2>&1 | grep -v lgb_init
done
} > gpurun_out/r2_v26_stack_limit.txt 2>&1
cat gpurun_out/r2_v26_stack_limit.txt | cut -c1-200
