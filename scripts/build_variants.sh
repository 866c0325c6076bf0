#!/bin/bash
# builds kernel variants for experiments: ./build_variants.sh "NAME:-DFLAG=.. -DFLAG2=.." ...
set -e
cd "$(dirname "$0")/.."
mkdir -p build
for spec in "$@"; do
  name="${spec%%:*}"; flags="${spec#*:}"
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false $flags \
    -Xcompiler -fPIC,-ffp-contract=off -shared -o build/lib_$name.so \
    lasgun_b200/csrc/lgb_kernels.cu lasgun_b200/csrc/lgb_gpubuild.cu lasgun_b200/csrc/lgb_grid.cu lasgun_b200/csrc/lgb_api.cu lasgun_b200/csrc/lgb_build.cpp lasgun_b200/csrc/lgb_parallel.cpp lasgun_b200/csrc/host/lasgun_host.cpp &
done
wait
ls build/
