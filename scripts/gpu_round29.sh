#!/bin/bash
# k_cprimary's spread: order of allocations (scene first / film first) x register budget (8 / 7 / 6 blocks of 128 threads per SM)
mkdir -p gpurun_out
{
for so in "" build/lib_cp7.so build/lib_cp6.so; do
  echo "== ${so:-base}: profile_kernels (scene first)"; LASGUN_B200_SO=${so:+$PWD/$so} python scripts/profile_kernels.py mixed4k | cut -c1-60
  echo "== ${so:-base}: film first"; LASGUN_B200_SO=${so:+$PWD/$so} TRIALS=2 python scripts/diag_ctx_placement.py
done
} > gpurun_out/r2_v32_cprimary_regs.txt 2>&1
cat gpurun_out/r2_v32_cprimary_regs.txt
