#!/usr/bin/env python3
"""ncu report -> transposed raw page (one row per metric, one column per launch): profiles/*_raw.csv.

    python scripts/ncu_raw_transposed.py gpurun_out/prof.ncu-rep profiles/out_raw.csv
"""
import csv
import subprocess
import sys

src, dst = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
with open(dst, "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["metric", "unit"] + [f"launch{i}" for i in range(len(data))])
    for c, name in enumerate(hdr):
        w.writerow([name, units[c]] + [r[c] for r in data])
print(f"{dst}: {len(hdr)} metrics x {len(data)} launches")
