#!/usr/bin/env python
"""Per-CUDA-source-line stall-sample summary of an .ncu-rep captured with --import-source on.
  python scripts/ncu_lines.py gpurun_out/x.ncu-rep [top_n]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Line No"][0]
hdr, data = rows[hi], rows[hi + 1:]
si, ie = hdr.index("# Samples"), hdr.index("Instructions Executed")
agg = [(int(r[0]), int(r[si]), int(r[ie]), r[1].strip()[:105]) for r in data if r[0].isdigit() and r[2] == '-']
tot = sum(a[1] for a in agg); ti = sum(a[2] for a in agg)
print(f"# total samples {tot}, warp instructions {ti}")
for a in sorted(agg, key=lambda a: -a[1])[:top]:
    print(f"{a[0]:5d} {a[1]:7d} {100*a[1]/tot:5.1f}% {a[2]:10d}  {a[3]}")
