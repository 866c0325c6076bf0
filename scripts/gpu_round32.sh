#!/bin/bash
# fast / slow order under ncu: does the slow process read more DRAM (L2 conflicts) or just wait longer?
mkdir -p gpurun_out
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct,lts__t_sectors_op_read.sum,lts__t_sectors_op_write.sum,lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum,lts__t_sectors_srcunit_tex_op_read_lookup_hit.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,lts__t_sectors_srcnode_gpc_aperture_peer.sum,lts__average_t_sector_hit_rate_realtime.pct,l1tex__m_xbar2l1tex_read_sectors.sum"
for m in scene_first film_first; do
  MODE=$m ncu --clock-control none --metrics $M -k regex:k_cprimary -s 2 -c 1 --csv --log-file gpurun_out/ncu_order_$m.csv python scripts/diag_order.py > gpurun_out/ncu_order_$m.log 2>&1
  echo "== $m rc=$?"; python - <<PY
import csv
rows=list(csv.reader(l for l in open("gpurun_out/ncu_order_$m.csv") if l.startswith('"')))
h=rows[0]; iN=h.index("Metric Name"); iV=h.index("Metric Value"); iU=h.index("Metric Unit")
for r in rows[1:]: print(f"  {r[iN]:70s} {r[iV]:>20s} {r[iU]}")
PY
done
