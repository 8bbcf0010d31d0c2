"""Summarise an ncu launch list (`--metrics gpu__time_duration.sum --csv --log-file X`) by kernel: share of the summed
device time, launches, average duration.  Usage: python tools/ncu_launch_summary.py launches.csv "<command line>" """
import csv
import sys
from collections import defaultdict

path, cmd = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
rows = []
with open(path, newline="") as f:
    lines = [ln for ln in f if ln.startswith('"')]
rd = csv.reader(lines)
hdr = next(rd)
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}
agg = defaultdict(lambda: [0, 0.0])
for r in rd:
    if len(r) <= iv:
        continue
    try:
        us = float(r[iv].replace(",", "")) * scale.get(r[iu], 1e-3)
    except ValueError:
        continue
    name = r[ik].split("(")[0][:110]
    agg[name][0] += 1
    agg[name][1] += us
tot = sum(v[1] for v in agg.values())
n = sum(v[0] for v in agg.values())
ours = sum(v[1] for k, v in agg.items() if k.startswith(("ga::", "void ga::", "sa::", "void sa::", "tc::", "void tc::")))
print("# ncu launch list (gpu__time_duration.sum, --clock-control none) of")
print(f"#   {cmd}")
print("# per-launch times are cold-cache and serialised: compare SHARES.")
print(f"total {tot / 1e3:.1f} ms over {n} launches; this repo's kernels: {100 * ours / tot:.1f} % of the device time")
print("# share%  launches  avg_us  kernel")
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:60]:
    print(f"{100 * t / tot:6.2f} {c:7d} {t / c:8.1f}  {k}")
