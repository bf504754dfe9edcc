#!/usr/bin/env python
"""Aggregate stall reasons / opcode mix of one kernel from an .ncu-rep: python tools/ncu_stalls.py rep <kernel regex> [skip]"""
import csv, subprocess, sys, io, collections
rep, rx = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{rx}", "--launch-skip", skip, "--launch-count", "1"],
                     capture_output=True, text=True).stdout
lines = out.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rows = [r for r in csv.DictReader(io.StringIO("\n".join(lines[start:]))) if (r["# Samples"] or "0").isdigit()]
rows = rows[: len(rows) // 2] if len(rows) > 2 and rows[0]["Source"] == rows[len(rows) // 2]["Source"] else rows
tot = sum(int(r["# Samples"] or 0) for r in rows)
stall_cols = [c for c in rows[0].keys() if c.startswith("stall_") and "Not Issued" not in c]
agg = collections.Counter()
for r in rows:
    for c in stall_cols:
        agg[c] += int(r[c] or 0)
print("total samples", tot)
for c, v in agg.most_common(12):
    print(f"  {c:28s} {v:7d} {100*v/max(tot,1):5.1f}%")
ops = collections.Counter(); ex = collections.Counter()
exec_col = "# Warp Instructions Executed" if "# Warp Instructions Executed" in rows[0] else None
for r in rows:
    op = r["Source"].strip().split()
    op = [o for o in op if not o.startswith("@")]
    name = op[0].split(".")[0] if op else "?"
    ops[name] += int(r["# Samples"] or 0)
    if exec_col: ex[name] += int(r[exec_col] or 0)
print("by opcode (samples | warp-instructions executed):")
for n, v in ops.most_common(25):
    print(f"  {n:14s} {v:7d} {100*v/max(tot,1):5.1f}%   {ex[n]:12d}")
if exec_col: print("total warp instructions", sum(ex.values()))
