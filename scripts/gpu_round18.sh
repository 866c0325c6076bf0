#!/bin/bash
# grid walks split into filter runs + convergent exact tests (LGB_WALK_SPLIT): parity, then A/B against the fused walk and two occupancy variants
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
{
for so in "" build/lib_nosplit.so build/lib_split_g3.so build/lib_split_c6.so; do
  echo "== ${so:-in-tree (split)}"
  LASGUN_B200_SO=${so:+$PWD/$so} python scripts/profile_kernels.py mixed4k | cut -c1-150
  LASGUN_B200_SO=${so:+$PWD/$so} python scripts/profile_frame.py spheres1m 4 | tail -1 | cut -c1-70
  LASGUN_B200_SO=${so:+$PWD/$so} python scripts/profile_frame.py mesh1m 4 | tail -1 | cut -c1-70
done
} > gpurun_out/r2_v23_walk_split_ab.txt 2>&1
cat gpurun_out/r2_v23_walk_split_ab.txt
python bench.py --steps 5 --warmup 3 --no-other-configs > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; python - <<PY
import json
d=json.loads(open("gpurun_out/bench_n1.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","kernel_ms_per_frame")}); print(json.dumps(d["e2e"])[:900])
PY
