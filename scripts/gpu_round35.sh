#!/bin/bash
mkdir -p gpurun_out
{
for r in 0 2048; do for m in scene_first raw64kb_after_scene film_first small_torch_first raw_film_first; do echo -n "DevBuf rounding $r KB: "; LGB_DEVBUF_ROUND_KB=$r MODE=$m python scripts/diag_order.py 2>&1 | tail -1; done; done
} > gpurun_out/r2_v37_devbuf_round.txt 2>&1
cat gpurun_out/r2_v37_devbuf_round.txt
