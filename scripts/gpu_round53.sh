#!/bin/bash
# DRAM traffic of the frame's kernels for the final sources (bench.py quotes it only for a build of the same hash)
mkdir -p gpurun_out
python scripts/ncu_traffic.py mixed4k r2_v56 2>&1 | tail -6
python scripts/ncu_summary.py gpurun_out/ncu_r2_v56.ncu-rep > gpurun_out/r2_v56_ncu_summary.txt 2>&1
cp gpurun_out/ncu_traffic.json profiles/ncu_traffic.json
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-other-configs > gpurun_out/r2_v56_bench_check.json 2>/dev/null
python - <<PY
import json
d=json.loads(open("gpurun_out/r2_v56_bench_check.json").read().strip().splitlines()[-1]); r=d["roofline"]
print(d["ms_per_step"], r["kernel"], r["frac"], r["traffic"])
PY
