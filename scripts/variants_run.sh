#!/bin/bash
# runs profile_frame for the in-tree library and every build/lib_*.so variant
w=${1:-mixed4k}
echo "== base"; python scripts/profile_frame.py $w 3 | tail -1 | cut -c1-60
for so in build/lib_*.so; do echo "== $so"; LASGUN_B200_SO=$PWD/$so python scripts/profile_frame.py $w 3 | tail -1 | cut -c1-60; done
