#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
python scripts/diag_group_pool.py 2>&1 | tail -2
for w in spheres1m mesh1m mixed4k; do LGB_TIMING=1 python scripts/e2e_breakdown.py $w 2>&1 | tail -24; done > gpurun_out/e2e_breakdown_timing.txt; grep -E "it3|light grids\] [0-9]|sort|camera grid" gpurun_out/e2e_breakdown_timing.txt | cut -c1-220 | tail -30
LGB_TIMING=1 LGB_GPUBUILD_DEBUG=1 python scripts/e2e_breakdown.py mesh1m 2>&1 | grep gpu_build | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 2 --steps 5 --warmup 3 --no-other-configs > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench rc=$?"; python - <<PY
import json
d=json.loads(open("gpurun_out/bench_n2.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","kernel_ms_per_frame")}, d["config"]["film_gather"]); print(json.dumps(d["e2e"], indent=1)[:1800])
PY
