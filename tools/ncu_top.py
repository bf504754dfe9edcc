#!/usr/bin/env python
"""Top stall-sample SASS lines of one kernel from an .ncu-rep:  python tools/ncu_top.py rep.ncu-rep <kernel regex> [skip] [n]"""
import csv, subprocess, sys, io
rep, rx = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
n = int(sys.argv[4]) if len(sys.argv) > 4 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{rx}", "--launch-skip", skip, "--launch-count", "1"],
                     capture_output=True, text=True).stdout
lines = out.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
print(lines[start - 1][:160])
rows = [r for r in csv.DictReader(io.StringIO("\n".join(lines[start:]))) if (r["# Samples"] or "0").isdigit()]
tot = sum(int(r["# Samples"] or 0) for r in rows)
print("total samples", tot, "instructions", len(rows))
stall_cols = [c for c in rows[0].keys() if c.startswith("stall_")]
for i, r in enumerate(rows):
    r["_i"] = i
top = sorted(rows, key=lambda r: -int(r["# Samples"] or 0))[:n]
for r in top:
    st = sorted(((int(r[c] or 0), c) for c in stall_cols), reverse=True)[:2]
    print(f'{int(r["# Samples"]):6d} {100*int(r["# Samples"])/max(tot,1):5.1f}%  #{r["_i"]:4d} {r["Source"].strip()[:90]:90s} {st}')
