#!/bin/bash
w=${1:-mixed4k}
echo "== base"; python scripts/e2e_breakdown.py $w | tail -1
for so in build/lib_*.so; do echo "== $so"; LASGUN_B200_SO=$PWD/$so python scripts/e2e_breakdown.py $w | tail -1; done
