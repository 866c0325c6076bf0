#!/bin/bash
# The BVH traversal kernels (k_primary, k_shadow) with the north star's two untried prescriptions, per-ray walk forced
# (grids and beams off): stacks in shared memory, top levels of the BVH in shared memory.  Prints ms/frame per variant and workload.
mkdir -p gpurun_out
export LGB_CAMERA_GRID=0 LGB_LIGHT_GRIDS=0 LGB_BEAMS=0
for w in mesh1m spheres1m cornell mixed4k; do
  echo "== $w base"; python scripts/profile_frame.py $w 3 | tail -1 | cut -c1-60
  for so in build/lib_*.so; do echo "== $w $so"; LASGUN_B200_SO=$PWD/$so python scripts/profile_frame.py $w 3 | tail -1 | cut -c1-60; done
done 2>&1 | tee gpurun_out/bvh_variants_${TAG:-r2}.txt
# ncu: local-memory traffic and L1 pressure of k_primary, base vs shared-memory stack vs staged top levels (mesh1m: one k_primary launch)
for v in base smemstack smemtop256; do
  so=""; [ $v != base ] && so="LASGUN_B200_SO=$PWD/build/lib_$v.so"
  env $so ncu --metrics gpu__time_duration.sum,smsp__sass_inst_executed_op_local_ld.sum,smsp__sass_inst_executed_op_local_st.sum,smsp__sass_inst_executed_op_shared_ld.sum,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,l1tex__t_sector_hit_rate.pct \
    --clock-control none -k "regex:k_primary|k_shadow" -c 3 --csv --log-file gpurun_out/ncu_bvh_$v.csv python scripts/profile_frame.py mixed4k 1 > /dev/null 2>&1
  python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/ncu_bvh_$v.csv")) if len(r)>10 and r[0].isdigit()]
d={}
for r in rows: d.setdefault((r[0], r[4][:40]), {})[r[-3]]=r[-1]
for (i,k),m in d.items(): print("$v", k, {a.split("__")[-1][:34]: b for a,b in m.items()})
PY
done 2>&1 | tee -a gpurun_out/bvh_variants_${TAG:-r2}.txt
