"""Where do the warps of a multi-role kernel wait?  Lists the mbarrier try-wait / bar.sync sites of one kernel (SASS level,
from an `ncu --set full --import-source on` report) with their share of the stall samples, plus the busiest SASS rows.
The mbarrier's shared-memory offset in the operand identifies the barrier array entry, i.e. which role waits for what.
Usage: python tools/ncu_waits.py report.ncu-rep [min_share_percent]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
min_share = float(sys.argv[2]) if len(sys.argv) > 2 else 0.4
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "sass", "--csv"], capture_output=True,
                     text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = next(r for r in rows if r and r[0] == "Address")
isrc, ismp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
data = [(r[isrc].strip(), int(r[ismp] or 0), int(r[iex] or 0)) for r in rows[rows.index(hdr) + 1:] if len(r) > iex]
tot = sum(d[1] for d in data) or 1
print(f"{rows[0][1][:100]}\ntotal samples {tot}")
for i, (s, n, e) in enumerate(data):
    if ("SYNCS" in s and "TRYWAIT" in s) or s.startswith("BAR.") or "BAR.SYNC" in s:
        grp = sum(d[1] for d in data[i:i + 4])      # the try-wait and the branch / nanosleep that loop on it
        if 100.0 * grp / tot >= min_share:
            print(f"{100.0 * grp / tot:5.1f}% of samples  sass row {i:5d}  executed {e:9d}  {s[:70]}")

# samples per 100-row SASS region with the marker instructions found there (the roles of a warp-specialised kernel are
# distinct code regions: producer = UTMALDG, issuers = UTCHMMA, epilogue = LDTM + stores, softmax = LDTM + MUFU + STTM)
print("region share of samples / max executed / markers")
for b in range(0, len(data), 100):
    seg = data[b:b + 100]
    s = sum(d[1] for d in seg)
    if 100.0 * s / tot < 0.3:
        continue
    kinds = sorted({d[0].split()[0] for d in seg if any(m in d[0] for m in
                    ("TRYWAIT", "UTMA", "UTCHMMA", "LDTM", "STTM", "MUFU.EX2", "BAR.SYNC", "STG.E.128", "UBLKCP"))})
    print(f"rows {b:5d}+  {100.0 * s / tot:5.1f}%  {max(d[2] for d in seg):9d}  {kinds}")
