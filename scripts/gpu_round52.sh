#!/bin/bash
# the BVH traversal kernels (k_primary, k_shadow: persistent, 128 threads) at 8 / 10 / 12 blocks per SM = 64 / 48 / 40 registers
mkdir -p gpurun_out
{
for v in "" trav10 trav12; do
  so=${v:+$PWD/build/lib_$v.so}
  echo "== ${v:-base (8 blocks)}"
  for w in simple cornell mesh1m; do LASGUN_B200_SO=$so python scripts/profile_kernels.py $w | cut -c1-60 | tail -1; done
  LASGUN_B200_SO=$so LGB_LIGHT_GRIDS=0 LGB_CAMERA_GRID=0 python scripts/profile_kernels.py mixed4k | cut -c1-60 | tail -1
done
} > gpurun_out/r2_v54_traversal_occupancy.txt 2>&1
cat gpurun_out/r2_v54_traversal_occupancy.txt
