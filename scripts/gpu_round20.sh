#!/bin/bash
# k_cprimary<SETUP> without the per-thread copy of the kernel's DevScene parameter to the stack
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
{
for i in 1 2 3; do python scripts/profile_kernels.py mixed4k | cut -c1-60; done
DIAG_STEPS=first_on_tstream,events python scripts/diag_bench_gap.py
python scripts/profile_kernels.py cornell | cut -c1-60
python scripts/profile_kernels.py simple | cut -c1-60
} > gpurun_out/r2_v24_no_param_copy.txt 2>&1
cat gpurun_out/r2_v24_no_param_copy.txt
python bench.py --steps 5 --warmup 3 --no-other-configs --no-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; python - <<PY
import json
d=json.loads(open("gpurun_out/bench_n1.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","kernel_ms_per_frame")}); print(json.dumps(d["e2e"])[:900])
PY
