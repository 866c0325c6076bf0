#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
{
for so in "" build/lib_sh64.so build/lib_gs32.so; do
  echo "== ${so:-base}"; LASGUN_B200_SO=${so:+$PWD/$so} python scripts/profile_kernels.py mixed4k | cut -c1-60
  LASGUN_B200_SO=${so:+$PWD/$so} python scripts/profile_kernels.py spheres1m | cut -c1-60 | tail -1
  LASGUN_B200_SO=${so:+$PWD/$so} python scripts/profile_kernels.py cornell | cut -c1-60 | tail -1
  LASGUN_B200_SO=${so:+$PWD/$so} python scripts/profile_kernels.py simple | cut -c1-60 | tail -1
done
} > gpurun_out/r2_v43_block_sizes.txt 2>&1
cat gpurun_out/r2_v43_block_sizes.txt
