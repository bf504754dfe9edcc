#!/usr/bin/env python
"""DRAM traffic per launch of each kernel in an .ncu-rep (dram__bytes_read.sum + dram__bytes_write.sum, averaged over the
captured launches): python tools/ncu_traffic.py rep.ncu-rep > profiles/rNN_ncu_traffic.json"""
import collections, csv, io, json, subprocess, sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
TIME = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
for r in data:
    name = r[col["Kernel Name"]].split("(")[0].replace("void ", "").split("<")[0].replace("b200x::", "")
    rd = float(r[col["dram__bytes_read.sum"]]) * UNIT[units[col["dram__bytes_read.sum"]]]
    wr = float(r[col["dram__bytes_write.sum"]]) * UNIT[units[col["dram__bytes_write.sum"]]]
    ms = float(r[col["gpu__time_duration.sum"]]) * TIME[units[col["gpu__time_duration.sum"]]]
    a = agg[name]
    a[0] += 1; a[1] += rd; a[2] += wr; a[3] += ms
print(json.dumps({k: {"launches": v[0], "dram_read_bytes_per_launch": v[1] / v[0], "dram_write_bytes_per_launch": v[2] / v[0],
                      "traffic_bytes_per_launch": (v[1] + v[2]) / v[0], "ncu_ms_per_launch": v[3] / v[0]} for k, v in agg.items()}, indent=1))
