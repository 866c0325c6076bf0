#!/bin/bash
mkdir -p gpurun_out
{
for nc in "" 1; do for m in scene_first film_first raw64kb_after_scene; do echo -n "LGB_NO_COUNTERS=$nc: "; LGB_DEVBUF_ROUND_KB=0 LGB_NO_COUNTERS=$nc MODE=$m python scripts/diag_order.py 2>&1 | tail -1; done; done
} > gpurun_out/r2_v38_counters.txt 2>&1
cat gpurun_out/r2_v38_counters.txt
