"""Per-kernel summary of an `ncu --set full` report: python scripts/ncu_summary.py <file.ncu-rep>  (reads `ncu -i ... --page raw --csv`)."""
import csv, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h, units = rows[0], rows[1]
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__warps_eligible.avg.per_cycle_active",
        "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum", "smsp__sass_inst_executed_op_global_ld.sum"]
ix = {k: h.index(k) for k in KEYS if k in h}
kn = h.index("Kernel Name")
print("# source:", rep)
for r in rows[2:]:
    print("\n==", r[kn][:110])
    for k, i in ix.items():
        print(f"  {k} = {r[i]} {units[i]}")
    st = [(float(r[i] or 0), k.replace("smsp__pcsamp_warps_issue_stalled_", "")) for i, k in enumerate(h) if "pcsamp_warps_issue_stalled" in k and "not_issued" not in k]
    tot = sum(v for v, _ in st) or 1
    print("  stall samples:", ", ".join(f"{k} {100 * v / tot:.0f}%" for v, k in sorted(st, reverse=True)[:7]))
