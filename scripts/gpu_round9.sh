#!/bin/bash
mkdir -p gpurun_out
for cfg in "1.0 1024" "0.5 1024" "0.5 2048" "0.25 2048" "2.0 1024"; do set -- $cfg; echo "== LGB_GRID_CELL=$1 LGB_GRID_RES=$2"; for w in mixed4k spheres1m; do LGB_TIMING=1 LGB_GRID_CELL=$1 LGB_GRID_RES=$2 python scripts/profile_frame.py $w 3 2>&1 | grep -E "frame 2|^\[light grids\] " | cut -c1-110; done; done 2>&1 | tee gpurun_out/grid_cell_${TAG:-r2}.txt
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_${TAG:-r2}.json 2> gpurun_out/bench_${TAG:-r2}.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_${TAG:-r2}.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_${TAG:-r2}.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","kernel_ms_per_frame","gpu_launches")}); print(json.dumps(d["e2e"])[:600])
r=d["roofline"]; print({k:r[k] for k in ("kernel","share_of_frame","bound","achieved","peak","frac","traffic","launch_ms")})
for k in r["kernels"]: print(k["kernel"], round(k["launch_ms"],3), round(k["frac"],3), k["bound"])
print(d.get("cpu_baseline")); 
for n,q in d.get("other_configs",{}).items(): print(n, {k:q.get(k) for k in ("kernel_ms_per_frame","value","e2e_ms_per_frame","dominant_kernel","error")})
PY
