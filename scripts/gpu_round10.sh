#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
for cfg in "-1 1024" "1.0 1024" "0.5 1024" "0.5 2048" "-1 2048"; do set -- $cfg; echo "== LGB_GRID_CELL=$1 LGB_GRID_RES=$2"; for w in mixed4k spheres1m mesh1m; do LGB_TIMING=1 LGB_GRID_CELL=$1 LGB_GRID_RES=$2 python scripts/profile_frame.py $w 3 2>&1 | grep -E "frame 2|^\[light grids\] [0-9]|camera grid" | cut -c1-120; done; done 2>&1 | tee gpurun_out/grid_cell_${TAG:-r2}.txt
