#!/bin/bash
mkdir -p gpurun_out
{
for v in s0b8 s0b12 s0b14 s0t64b20 s0t64b24; do
  so=build/lib_$v.so
  echo "== $v"; LASGUN_B200_SO=$PWD/$so python scripts/profile_kernels.py mixed4k | cut -c1-60 | head -1
  LASGUN_B200_SO=$PWD/$so python scripts/profile_kernels.py spheres1m | cut -c1-60 | head -1
done
} > gpurun_out/r2_v48_cprimary_nostage.txt 2>&1
cat gpurun_out/r2_v48_cprimary_nostage.txt
