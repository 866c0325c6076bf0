#!/bin/bash
mkdir -p gpurun_out
{
DIAG_STEPS=plain,tstream,sampler,events,nodecount python scripts/diag_bench_gap.py
DIAG_STEPS=first_on_tstream,sampler,events python scripts/diag_bench_gap.py
} > gpurun_out/diag_bench_gap.txt 2>&1
cat gpurun_out/diag_bench_gap.txt | cut -c1-250
