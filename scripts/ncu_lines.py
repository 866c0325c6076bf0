"""Aggregates `ncu --page source --csv --print-source sass,cuda` output by CUDA source line."""
import collections
import csv
import sys

def toint(s):
    try:
        return int(float(s))
    except Exception:
        return 0

rows = list(csv.reader(open(sys.argv[1])))
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hi = [i for i, r in enumerate(rows) if r and r[0] == "Line No"]
inst, samp, thr = collections.Counter(), collections.Counter(), collections.Counter()
for si, h in enumerate(hi):
    hdr = rows[h]
    end = hi[si + 1] - 2 if si + 1 < len(hi) else len(rows)
    fpath = rows[h - 2][1] if h >= 2 and rows[h - 2] and rows[h - 2][0] == "File Path" else "?"
    col = {}
    for i, n in enumerate(hdr):
        col.setdefault(n, i)
    for r in rows[h + 1:end]:
        if len(r) < len(hdr) or not r[0].strip().isdigit():
            continue
        key = (fpath.split("/")[-1], int(r[0]), r[1].strip()[:100])
        inst[key] += toint(r[col["Instructions Executed"]]); samp[key] += toint(r[col["# Samples"]]); thr[key] += toint(r[col["Thread Instructions Executed"]])
T, S = sum(inst.values()) or 1, sum(samp.values()) or 1
print("total warp instr", T, "samples", S)
for k, v in sorted(inst.items(), key=lambda kv: -samp[kv[0]])[:topn]:
    print(f"{samp[k] / S * 100:5.1f}% samp {v / T * 100:5.1f}% inst  thr/inst {thr[k] / max(v, 1):4.1f}  {k[0]}:{k[1]}  {k[2]}")
