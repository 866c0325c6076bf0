#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
{
echo "== base (fused light-grid walk, 40 registers)"; python scripts/profile_kernels.py mixed4k | cut -c1-60; python scripts/profile_kernels.py spheres1m | cut -c1-60 | tail -4
for v in gs20 gs16; do
  so=build/lib_$v.so
  echo "== $v"; LASGUN_B200_SO=$PWD/$so python scripts/profile_kernels.py mixed4k | cut -c1-60 | grep gshadow
  LASGUN_B200_SO=$PWD/$so python scripts/profile_kernels.py spheres1m | cut -c1-60 | grep gshadow
done
} > gpurun_out/r2_v51_gshadow_fused.txt 2>&1
cat gpurun_out/r2_v51_gshadow_fused.txt
