"""Opcode histogram of one kernel from `ncu --page source --csv --print-source sass,cuda` output: python scripts/ncu_opcodes.py file.csv warps"""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
warps = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
cnt, samp, tot = collections.Counter(), collections.Counter(), 0
hdr = None
for r in rows:
    if r and r[0] == "Line No":
        hdr = {n: i for i, n in reversed(list(enumerate(r)))}
        continue
    if not hdr or len(r) < 10 or r[0] != "" or not r[2].startswith("0x"):
        continue
    sass = r[3].split()
    if not sass:
        continue
    op = sass[1] if sass[0].startswith("@") and len(sass) > 1 else sass[0]
    op = ".".join(op.split(".")[:2]) if op.startswith(("MUFU", "LDG", "STG", "LDL", "STL", "LDS", "STS", "ATOM", "RED")) else op.split(".")[0]
    n = int(float(r[7] or 0))
    cnt[op] += n; samp[op] += int(float(r[6] or 0)); tot += n
S = sum(samp.values()) or 1
print("total warp instr", tot, "per warp", round(tot / warps, 1))
for k, v in cnt.most_common(40):
    print(f"{k:14s} {v / tot * 100:5.1f}% inst {samp[k] / S * 100:5.1f}% samp  {v / warps:8.1f} per warp")
