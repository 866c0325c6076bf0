#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
for w in mixed4k spheres1m mesh1m; do LGB_TIMING=1 python scripts/e2e_breakdown.py $w 2>&1 | grep -E " e2e |ensure_bvh" | tail -2; done | tee gpurun_out/e2e_breakdown_${TAG:-r2}.txt
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_${TAG:-r2}.json 2> gpurun_out/bench_${TAG:-r2}.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_${TAG:-r2}.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_${TAG:-r2}.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","kernel_ms_per_frame","gpu_launches")}); print(json.dumps(d["e2e"])[:700])
r=d["roofline"]; print({k:r[k] for k in ("kernel","share_of_frame","bound","achieved","peak","frac","traffic","launch_ms")})
for k in r["kernels"]: print(k["kernel"], round(k["launch_ms"],3), round(k["frac"],3), k["bound"])
print(d.get("cpu_baseline"));
for n,q in d.get("other_configs",{}).items(): print(n, {k:q.get(k) for k in ("kernel_ms_per_frame","value","e2e_ms_per_frame","error")}, (q.get("dominant_kernel") or {}).get("kernel"), (q.get("dominant_kernel") or {}).get("frac"))
PY
python bench.py --impl reference --steps 2 --warmup 1 --ref-seconds 30 > gpurun_out/bench_ref_${TAG:-r2}.json 2>&1; tail -c 600 gpurun_out/bench_ref_${TAG:-r2}.json
