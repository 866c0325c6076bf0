#!/bin/bash
mkdir -p gpurun_out
{
for v in stage0 gst128; do
  so=build/lib_$v.so
  echo "== $v"; LASGUN_B200_SO=$PWD/$so python scripts/profile_kernels.py mixed4k | cut -c1-60
  LASGUN_B200_SO=$PWD/$so python scripts/profile_kernels.py spheres1m | cut -c1-60 | tail -4
done
} > gpurun_out/r2_v47_stage.txt 2>&1
cat gpurun_out/r2_v47_stage.txt
