#!/bin/bash
mkdir -p gpurun_out
{
for v in split0 split0pf setupf; do
  so=build/lib_$v.so
  [ -f $so ] || continue
  echo "== $v"; LASGUN_B200_SO=$PWD/$so python scripts/profile_kernels.py mixed4k | cut -c1-60
  LASGUN_B200_SO=$PWD/$so python scripts/profile_kernels.py spheres1m | cut -c1-60 | tail -1
done
} > gpurun_out/r2_v50_walk_variants.txt 2>&1
cat gpurun_out/r2_v50_walk_variants.txt
