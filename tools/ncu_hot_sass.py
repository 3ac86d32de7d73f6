#!/usr/bin/env python
"""Top stall sites of one kernel launch from an .ncu-rep source page (SASS level).
   usage: python tools/ncu_hot_sass.py rep.ncu-rep <launch index> [top N]"""
import csv, io, subprocess, sys
rep, idx = sys.argv[1], int(sys.argv[2]); top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(idx), "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
print(rows[0][1][:150])
hdr = rows[1]; col = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[col["# Samples"]]) for r in data)
print("total samples", tot, "instructions", len(data))
agg = {s: sum(int(r[col[s]]) for r in data) for s in stalls}
print("stall totals:", {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v > tot * 0.01})
order = sorted(range(len(data)), key=lambda i: -int(data[i][col["# Samples"]]))[:top]
for i in sorted(order):
    r = data[i]
    n = int(r[col["# Samples"]])
    dom = sorted(((int(r[col[s]]), s) for s in stalls), reverse=True)[:2]
    print(f"{i:5d} {n:6d} {100.0*n/tot:5.1f}%  {r[col['Source']].strip()[:70]:70s} {dom[0][1]}:{dom[0][0]} {dom[1][1]}:{dom[1][0]}  exec={r[col['Instructions Executed']]}")
# ---- samples grouped by execution count (identifies loop bodies: per chunk drain / per epilogue block / per tile ...)
from collections import defaultdict
g = defaultdict(lambda: [0, 0])
for r in data:
    e = int(r[col["Instructions Executed"]]); g[e][0] += int(r[col["# Samples"]]); g[e][1] += 1
print("samples by exec count (exec: samples, #instructions):")
for e, (s_, n_) in sorted(g.items(), key=lambda kv: -kv[1][0])[:14]:
    print(f"   exec={e:9d}: {s_:6d} samples ({100.0*s_/tot:4.1f}%) over {n_} instructions")
