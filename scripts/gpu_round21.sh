#!/bin/bash
# (1) k_cprimary with the tile's list staged per warp in shared memory, A/B; (2) why a process that lets torch allocate before the scene runs k_cprimary 1.5 ms slower
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
{
for so in "" build/lib_nostage.so build/lib_stage_c6.so; do
  echo "== ${so:-in-tree (stage)}"
  LASGUN_B200_SO=${so:+$PWD/$so} python scripts/profile_kernels.py mixed4k | cut -c1-60 | head -3
  LASGUN_B200_SO=${so:+$PWD/$so} DIAG_STEPS=first_on_tstream python scripts/diag_bench_gap.py
  LASGUN_B200_SO=${so:+$PWD/$so} python scripts/profile_frame.py spheres1m 4 | tail -1 | cut -c1-70
done
echo "== placement (in-tree)"
DIAG_STEPS=scene_first,tstream python scripts/diag_bench_gap.py
DIAG_STEPS=plain,tstream python scripts/diag_bench_gap.py
} > gpurun_out/r2_v25_stage_ab.txt 2>&1
cat gpurun_out/r2_v25_stage_ab.txt | cut -c1-200
