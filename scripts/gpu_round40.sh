#!/bin/bash
# k_gshadow in blocks of 256 / 128 / 64 threads (same 1024 threads per SM)
mkdir -p gpurun_out
{
for so in "" build/lib_gs128.so build/lib_gs64.so; do
  echo "== ${so:-base 256}"; LASGUN_B200_SO=${so:+$PWD/$so} python scripts/profile_kernels.py mixed4k | cut -c1-60
  LASGUN_B200_SO=${so:+$PWD/$so} python scripts/profile_kernels.py spheres1m | cut -c1-60 | grep gshadow
  LASGUN_B200_SO=${so:+$PWD/$so} python scripts/profile_kernels.py mesh1m | cut -c1-60 | grep gshadow
done
} > gpurun_out/r2_v41_gshadow_blocks.txt 2>&1
cat gpurun_out/r2_v41_gshadow_blocks.txt
