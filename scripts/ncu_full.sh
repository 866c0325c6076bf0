#!/bin/bash
# ncu --set full capture (with source) of the kernels matching a regex, one frame of a workload:
#   scripts/ncu_full.sh "k_setup|k_shade_lean" tag [workload] [count]
re=$1; tag=$2; w=${3:-mixed4k}; n=${4:-4}
mkdir -p gpurun_out
ncu --set full --import-source on --clock-control none -k "regex:$re" -c $n -o gpurun_out/ncu_$tag -f python scripts/profile_frame.py $w 1 > gpurun_out/ncu_$tag.log 2>&1
echo "ncu rc=$?"; ls -la gpurun_out/ncu_$tag.ncu-rep
