#!/bin/bash
mkdir -p gpurun_out
{ echo "== contexts in turn"; python scripts/diag_ctx_placement.py; echo "== with a growing dummy allocation between them"; SHIFT=1 python scripts/diag_ctx_placement.py; echo "== again, fresh process"; TRIALS=3 python scripts/diag_ctx_placement.py; } > gpurun_out/r2_v31_ctx_placement.txt 2>&1
cat gpurun_out/r2_v31_ctx_placement.txt
