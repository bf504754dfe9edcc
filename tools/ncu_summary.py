#!/usr/bin/env python
"""Compact per-launch summary of an .ncu-rep (ncu --set full):  python tools/ncu_summary.py rep.ncu-rep > profiles/xxx.md"""
import csv, io, subprocess, sys

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
WANT = [("gpu__time_duration.sum", "time"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor pipe %"),
        ("dram__bytes.sum.per_second", "dram B/s"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"), ("l1tex__m_xbar2l1tex_read_bytes.sum", "l2->sm"),
        ("sm__inst_issued.avg.pct_of_peak_sustained_active", "issue%"), ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma%"),
        ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu%"), ("launch__registers_per_thread", "regs"),
        ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("smsp__inst_executed.sum", "warp_inst")]
print("| kernel | " + " | ".join(n for _, n in WANT) + " |")
print("|---|" + "---|" * len(WANT))
for r in data:
    name = r[col["Kernel Name"]].split("(")[0].replace("void ", "")
    cells = []
    for m, _ in WANT:
        i = col.get(m)
        if i is None:
            cells.append("-")
            continue
        v = r[i]
        try:
            v = f"{float(v):.4g}"
        except ValueError:
            pass
        cells.append(f"{v} {units[i]}".strip())
    print(f"| {name} | " + " | ".join(cells) + " |")
