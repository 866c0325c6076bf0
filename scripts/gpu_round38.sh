#!/bin/bash
mkdir -p gpurun_out
LGB_TIMING=1 python bench.py --no-cpu-baseline > gpurun_out/bench_timing.json 2> gpurun_out/bench_timing.err
grep -n "spheres\|light grids\] [0-9]\|camera grid\|lgb_scene_create\|flatten\]" gpurun_out/bench_timing.err | tail -40
