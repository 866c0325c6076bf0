#!/bin/bash
# per-kernel launch list of one frame: ./scripts/launch_times.sh <workload> <tag>
w=${1:-mixed4k}; tag=${2:-tmp}
ncu --metrics gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k "regex:k_primary|k_setup|k_shadow|k_pretest|k_shade|k_resolve|k_beam|k_leafp|k_secondary|k_spawn|k_gather|k_sbeam|k_swalk" -c 80 --csv --log-file gpurun_out/launches_$tag.csv python scripts/profile_frame.py $w 1 > /dev/null 2>&1
python - <<PY
import csv, collections
rows=[r for r in csv.reader(open("gpurun_out/launches_$tag.csv")) if len(r)>10 and r[0].isdigit()]
d=collections.OrderedDict()
for r in rows:
    d.setdefault((r[0], r[4][:28]), {})[r[-3]]=r[-1]
tot=0
for (i,k),m in d.items():
    t=float(m["gpu__time_duration.sum"])/1e6; tot+=t
    print(f"{k:30s} {t:8.3f} ms  simt {m.get('smsp__thread_inst_executed_per_inst_executed.ratio','')}  issue {m.get('smsp__issue_active.avg.pct_of_peak_sustained_active','')}  winst {float(m.get('smsp__inst_executed.sum','0'))/1e9:.2f}G occ {m.get('sm__warps_active.avg.pct_of_peak_sustained_active','')}")
print("sum", round(tot,3))
PY
