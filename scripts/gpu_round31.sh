#!/bin/bash
mkdir -p gpurun_out
{
for m in scene_first film_first; do LGB_TIMING=1 MODE=$m python scripts/diag_order.py 2>&1 | grep -E "wave buffers|k_cprimary" | sort -u | tail -3; done
for al in 128 512 1024 2048; do for m in scene_first film_first raw_film_first; do echo "align $al MB:"; LGB_WAVE_ALIGN_MB=$al MODE=$m python scripts/diag_order.py 2>&1 | tail -1; done; done
for off in 2 4 8 16 32 64; do echo "align 1024 MB + $off MB:"; LGB_WAVE_ALIGN_MB=1024 LGB_WAVE_OFFSET_MB=$off MODE=scene_first python scripts/diag_order.py 2>&1 | tail -1; done
} > gpurun_out/r2_v34_wave_align.txt 2>&1
cat gpurun_out/r2_v34_wave_align.txt
