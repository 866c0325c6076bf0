#!/bin/bash
mkdir -p gpurun_out
for m in scene_first film_first torch_init_only small_torch_first raw_film_first scene_first_raw_film big_raw_first scene_first film_first; do MODE=$m python scripts/diag_order.py 2>&1 | tail -1; done > gpurun_out/r2_v33_order.txt
cat gpurun_out/r2_v33_order.txt
