#!/bin/bash
# The round's evidence in one GPU session: parity tests, the bench line, the ncu launch list of the bench command, a per-kernel
# `ncu --set full` summary with DRAM traffic stamped with the source hash (profiles/ncu_traffic.json), per-kernel live timings.
tag=${TAG:-r2_final}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python scripts/ncu_traffic.py mixed4k $tag 2>&1 | tail -8
python scripts/ncu_summary.py gpurun_out/ncu_$tag.ncu-rep > gpurun_out/${tag}_ncu_summary.txt 2>&1
cp gpurun_out/ncu_traffic.json profiles/ncu_traffic.json          # (so that the bench below quotes the capture of this very build)
python bench.py --steps 5 --warmup 3 > gpurun_out/${tag}_bench_mixed4k.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/${tag}_bench.err
python bench.py --impl reference --steps 2 --warmup 1 --ref-seconds 30 > gpurun_out/${tag}_reference_arm.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/${tag}_launches_bench.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-other-configs > /dev/null 2>&1; echo "ncu launches rc=$?"
python scripts/profile_kernels.py mixed4k > gpurun_out/${tag}_kernels.txt 2>&1; cat gpurun_out/${tag}_kernels.txt
python - <<PY
import json
d=json.loads(open("gpurun_out/${tag}_bench_mixed4k.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","kernel_ms_per_frame","gpu_launches")}); print(json.dumps(d["e2e"])[:800])
r=d["roofline"]; print({k:r[k] for k in ("kernel","share_of_frame","bound","achieved","peak","frac","traffic","launch_ms","traffic_source")})
for k in r["kernels"]: print(k["kernel"], round(k["launch_ms"],3), round(k["frac"],3), k["bound"], k.get("traffic"))
for n,q in d.get("other_configs",{}).items(): print(n, {k:q.get(k) for k in ("kernel_ms_per_frame","value","e2e_ms_per_frame","error")}, (q.get("dominant_kernel") or {}).get("kernel"), (q.get("dominant_kernel") or {}).get("frac"))
PY
