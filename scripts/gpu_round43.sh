#!/bin/bash
mkdir -p gpurun_out
{
for so in build/lib_gsb20.so build/lib_gsb24.so; do
  echo "== ${so:-base}"; LASGUN_B200_SO=${so:+$PWD/$so} python scripts/profile_kernels.py mixed4k | cut -c1-60
  LASGUN_B200_SO=${so:+$PWD/$so} python scripts/profile_kernels.py spheres1m | cut -c1-60 | grep gshadow
done
} > gpurun_out/r2_v45_gshadow_regs.txt 2>&1
cat gpurun_out/r2_v45_gshadow_regs.txt
