#!/bin/bash
mkdir -p gpurun_out
for m in scene_first raw2mb_before_init raw2mb_after_init raw2mb_after_scene raw64kb_after_scene two_scenes small_torch_then_big_kept; do MODE=$m python scripts/diag_order.py 2>&1 | tail -1; done > gpurun_out/r2_v36_order2.txt
cat gpurun_out/r2_v36_order2.txt
