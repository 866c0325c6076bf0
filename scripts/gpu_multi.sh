#!/bin/bash
# N-GPU session: all gpu tests (incl. the 2-GPU ones), the torchrun bench at N, a memcheck pass over the grid / band tests
N=${1:-2}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_n$N.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu_n$N.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench rc=$?"; tail -c 1500 gpurun_out/bench_n$N.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_n$N.json").read().strip().splitlines()[-1])
    print({k:d[k] for k in ("value","ms_per_step","kernel_ms_per_frame")}, d["config"]["film_gather"]); print(json.dumps(d["e2e"], indent=1)[:1800])
except Exception as e: print("no bench line", e)
PY
if [ "$2" = "memcheck" ]; then timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "light_grids or camera_grid or bands or lazy" > gpurun_out/memcheck.log 2>&1; echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|Invalid|passed|failed" gpurun_out/memcheck.log | tail -5; fi
