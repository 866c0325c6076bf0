#!/bin/bash
# block cache for the host arrays, builder top-level kernels, sub-tree nodes through scratch: parity + the full bench line
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; python - <<PY
import json
d=json.loads(open("gpurun_out/bench_n1.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","kernel_ms_per_frame")}); print(json.dumps(d["e2e"])[:900])
for k,v in d.get("other_configs",{}).items(): print(k, v.get("kernel_ms_per_frame"), v.get("e2e_ms_per_frame"))
PY
LGB_TIMING=1 python scripts/e2e_breakdown.py mesh1m 2>&1 | grep -E "it[123]|gpu_build" | tail -4
