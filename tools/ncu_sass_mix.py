#!/usr/bin/env python
"""Instruction mix + top stall sites of a kernel from `ncu --page source --csv`.
usage: tools/ncu_sass_mix.py rep.ncu-rep [topN]"""
import csv, io, subprocess, sys, collections
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
rows = list(csv.reader(io.StringIO(raw)))
start = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[start]; ix = {h: i for i, h in enumerate(hdr)}
cnt = collections.Counter(); samp = collections.Counter(); tot = 0; stot = 0
lines = []
for r in rows[start + 1:]:
    if len(r) < len(hdr): continue
    try: n = int(r[ix["Instructions Executed"]]); s = int(r[ix["# Samples"]])
    except ValueError: continue
    toks = r[ix["Source"]].split()
    op = toks[1] if toks[0].startswith("@") else toks[0]
    op = op.split(".")[0]
    cnt[op] += n; samp[op] += s; tot += n; stot += s
    lines.append((s, n, r[ix["Source"]].strip(), r[ix["stall_long_sb"]], r[ix["stall_wait"]], r[ix["stall_short_sb"]], r[ix["stall_math"]]))
print("instruction mix (executed warp-instructions), total %d" % tot)
for op, n in cnt.most_common(22):
    print(f"  {op:10s} {100*n/tot:5.1f}%   stall samples {100*samp[op]/max(stot,1):5.1f}%")
print("top stall sites")
for s, n, src, lsb, w, ssb, m in sorted(lines, reverse=True)[:top]:
    print(f"  {100*s/max(stot,1):5.1f}%  exec {n:9d}  long_sb {lsb:>6s} wait {w:>6s} short_sb {ssb:>5s} math {m:>5s}  {src}")
