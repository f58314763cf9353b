#!/usr/bin/env python3
"""Summarise an ncu report's SASS source page: stall reasons overall and per code region.
usage: ncu_regions.py report.ncu-rep [kernel-index] [region-size]"""
import collections, csv, subprocess, sys
rep = sys.argv[1]; kidx = int(sys.argv[2]) if len(sys.argv) > 2 else 0; step = int(sys.argv[3]) if len(sys.argv) > 3 else 200
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
names = [rows[i - 1][1] if i else "" for i in hi]
start = hi[kidx]; end = hi[kidx + 1] - 1 if kidx + 1 < len(hi) else len(rows)
hdr = rows[start]; ci = {h: i for i, h in enumerate(hdr)}
body = rows[start + 1:end]
tot = sum(int(r[ci["# Samples"]] or 0) for r in body)
print("kernel:", names[kidx][:100]); print("total samples", tot, "instrs", len(body))
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = collections.Counter()
reg = collections.OrderedDict()
for k, r in enumerate(body):
    b = k // step
    e = reg.setdefault(b, [0, collections.Counter(), 0, collections.Counter()])
    s = int(r[ci["# Samples"]] or 0)
    e[0] += s
    src = r[ci["Source"]].split()
    op = src[1] if src and src[0].startswith("@") else (src[0] if src else "")
    e[1][op.split(".")[0]] += s
    e[2] += int(r[ci["Instructions Executed"]] or 0)
    for h in stall_cols:
        v = r[ci[h]]
        if v:
            e[3][h] += int(v); agg[h] += int(v)
print("stalls:", [(k, f"{100*v/tot:.1f}%") for k, v in agg.most_common(10)])
for b, (s, c, ie, st) in reg.items():
    if s > tot * 0.015:
        print(f"@{b*step:6d} {100*s/tot:5.1f}%  exec={ie:9d}  ops={c.most_common(4)}  stalls={st.most_common(3)}")
