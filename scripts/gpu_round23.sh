#!/bin/bash
# wavefront entries stored / loaded with streaming hints: parity, then four processes each way (k_cprimary varies from process to process)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
{
for so in "" build/lib_nostream.so; do
  echo "== ${so:-in-tree (streaming hints)}"
  for i in 1 2 3 4; do
    LASGUN_B200_SO=${so:+$PWD/$so} python scripts/profile_kernels.py mixed4k | cut -c1-46 | tr '\n' '|'; echo
  done
  LASGUN_B200_SO=${so:+$PWD/$so} DIAG_STEPS=plain python scripts/diag_bench_gap.py
  LASGUN_B200_SO=${so:+$PWD/$so} DIAG_STEPS=first_on_tstream python scripts/diag_bench_gap.py
  LASGUN_B200_SO=${so:+$PWD/$so} python scripts/profile_frame.py spheres1m 4 | tail -1 | cut -c1-70
done
} > gpurun_out/r2_v27_wave_stream_ab.txt 2>&1
cat gpurun_out/r2_v27_wave_stream_ab.txt | cut -c1-250
