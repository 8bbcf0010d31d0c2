"""Per-source-line stall samples / instruction counts of one kernel from an `ncu --set full --import-source on` report.
Usage: python tools/ncu_lines.py report.ncu-rep [top_n]"""
import csv
import subprocess
import sys
from collections import defaultdict

rep = sys.argv[1]
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True,
                     text=True).stdout
rows = list(csv.reader(raw.splitlines()))
fname = ""
per = defaultdict(lambda: [0, 0, ""])
hdr = None
last_line = ""
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        isamp, iinst = hdr.index("# Samples"), hdr.index("Instructions Executed")
        continue
    if hdr is None or len(r) <= iinst:
        continue
    if r[0].strip():
        last_line = r[0]
        if len(r) > 2 and not r[2].strip():
            continue                 # the source line's own row repeats the sum of its SASS rows
    key = (fname, last_line)
    try:
        per[key][0] += int(r[isamp] or 0)
        per[key][1] += int(r[iinst] or 0)
    except ValueError:
        continue
    if r[1].strip():
        per[key][2] = r[1].strip()
tot = sum(v[0] for v in per.values()) or 1
toti = sum(v[1] for v in per.values()) or 1
print(f"total samples {tot}, warp instructions {toti}")
for (f, ln), (s, n, src) in sorted(per.items(), key=lambda kv: -kv[1][0])[:top_n]:
    print(f"{100 * s / tot:5.1f}% smp {100 * n / toti:5.1f}% inst  {f}:{ln:>5}  {src[:110]}")
