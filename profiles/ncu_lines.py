"""Aggregate an ncu report's source page by CUDA source line: python profiles/ncu_lines.py <report.ncu-rep> [top_n]."""
import csv
import subprocess
import sys
from collections import defaultdict

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
fname, hdr = None, None
lines = []
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        fname = r[1].split("/")[-1]
    elif len(r) > 5 and r[0] == "Line No":
        hdr = r
        ia = hdr.index("Instructions Executed")
        ism = hdr.index("# Samples")
    elif hdr and len(r) > ia and r[0].strip().isdigit():
        try:
            lines.append((fname, int(r[0]), r[1].strip()[:110], int(r[ia]), int(r[ism])))
        except ValueError:
            pass
tot = sum(x[3] for x in lines)
tots = sum(x[4] for x in lines)
print(f"total warp instructions {tot:.4g}, samples {tots}")
lines.sort(key=lambda x: -x[3])
acc = 0
for f, ln, src, n, s in lines[:top]:
    acc += n
    print(f"{100*n/tot:5.1f}% (cum {100*acc/tot:5.1f}%) smp {100*s/max(tots,1):5.1f}%  {f}:{ln}: {src}")
