"""DRAM traffic per launch of the frame's kernels from one `ncu --set full` capture -> gpurun_out/ncu_traffic.json, stamped with the hash
of the CUDA sources (bench.py quotes the file only for a build of the same sources).  Run on the GPU box:
    python scripts/ncu_traffic.py mixed4k r2_final"""
import csv, json, os, subprocess, sys
sys.path.insert(0, ".")
import bench
w = sys.argv[1] if len(sys.argv) > 1 else "mixed4k"
tag = sys.argv[2] if len(sys.argv) > 2 else "r2_final"
rep = f"gpurun_out/ncu_{tag}"
regex = "regex:k_cprimary|k_setup|k_gshadow|k_shade_lean|k_primary|k_beam|k_leafp|k_shadow|k_sbeam|k_swalk|k_pretest|k_shade"
r = subprocess.run(["ncu", "--set", "full", "--import-source", "on", "--clock-control", "none", "-k", regex, "-c", "12", "-o", rep, "-f",
                    sys.executable, "scripts/profile_frame.py", w, "1"], capture_output=True, text=True)
print("ncu rc", r.returncode, r.stdout[-300:], r.stderr[-300:])
raw = subprocess.run(["ncu", "-i", rep + ".ncu-rep", "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h = rows[0]
kn, rd, wr, tm = h.index("Kernel Name"), h.index("dram__bytes_read.sum"), h.index("dram__bytes_write.sum"), h.index("gpu__time_duration.sum")
units = rows[1]
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
out = {}
for row in rows[2:]:
    name = row[kn].split("<")[0].split("(")[0].replace("void ", "").strip()
    b = float(row[rd]) * scale[units[rd]] + float(row[wr]) * scale[units[wr]]
    out.setdefault(name, []).append(b)
    print(f"{name:20s} dram {b / 1e9:7.3f} GB  time {row[tm]} {units[tm]}")
traffic = {k: sum(v) / len(v) for k, v in out.items()}            # per launch (the per-light launches of one kernel averaged)
json.dump({"source_hash": bench.kernel_source_hash(), "capture": f"{tag}_ncu_summary.txt", "how": "ncu --set full --clock-control none, first frame of scripts/profile_frame.py",
           w: traffic}, open("gpurun_out/ncu_traffic.json", "w"), indent=1)
print(json.dumps(traffic))
