#!/bin/bash
# 2 GPUs: torchrun bench (device group after the memory-pool grant), then where capture's host time goes for the 1 M-primitive configs
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench rc=$?"; python - <<PY
import json
d=json.loads(open("gpurun_out/bench_n2.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","kernel_ms_per_frame")}, d["config"]["film_gather"]); print(json.dumps(d["e2e"], indent=1)[:1800])
PY
for w in spheres1m mesh1m; do LGB_TIMING=1 python scripts/e2e_breakdown.py $w 2>&1 | tail -40; done > gpurun_out/e2e_breakdown_timing.txt; tail -50 gpurun_out/e2e_breakdown_timing.txt | cut -c1-250
