#!/bin/bash
mkdir -p gpurun_out
for w in ${WORKLOADS:-mixed4k}; do scripts/variants_run.sh $w 2>&1 | tee -a gpurun_out/variants.log; done
LASGUN_B200_SO=$PWD/build/lib_nofuse.so scripts/ncu_full.sh "k_cprimary|k_gshadow|k_setup|k_shade_lean" ${TAG:-r2}_grids mixed4k 4
scripts/ncu_full.sh "k_surface" ${TAG:-r2}_surface mixed4k 1
