#!/bin/bash
# sliced shade + overlapped read-back (fixed), chunked scene upload, DevScene::self for the out-of-line instanced helpers: parity, kernels, full bench line
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
{
python scripts/profile_kernels.py mixed4k | cut -c1-60
python scripts/profile_kernels.py cornell | cut -c1-60
python scripts/profile_kernels.py simple | cut -c1-60
} > gpurun_out/r2_v28_kernels.txt 2>&1
cat gpurun_out/r2_v28_kernels.txt
python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; python - <<PY
import json
d=json.loads(open("gpurun_out/bench_n1.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","kernel_ms_per_frame")}); print(json.dumps(d["e2e"])[:900])
for k,v in d.get("other_configs",{}).items(): print(k, v.get("kernel_ms_per_frame"), v.get("e2e_ms_per_frame"))
PY
