#!/bin/bash
# striped work counters: parity, the allocation-order matrix again, the bench line
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
for m in scene_first film_first raw64kb_after_scene small_torch_first raw_film_first; do MODE=$m python scripts/diag_order.py 2>&1 | tail -1; done > gpurun_out/r2_v39_striped_counters.txt
cat gpurun_out/r2_v39_striped_counters.txt
python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; python - <<PY
import json
d=json.loads(open("gpurun_out/bench_n1.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","kernel_ms_per_frame")}); print(json.dumps(d["e2e"])[:900])
for k in d["roofline"]["kernels"]: print(k["kernel"], round(k["launch_ms"],3), round(k["frac"],3))
for k,v in d.get("other_configs",{}).items(): print(k, v.get("kernel_ms_per_frame"), v.get("e2e_ms_per_frame"))
PY
