#!/usr/bin/env python
"""Aggregate an ncu report's warp-stall samples by SOURCE LINE (needs -lineinfo; `--import-source on` not required for
file:line attribution).  python profiles/source_hotspots.py <report.ncu-rep> <kernel regex> [launch-skip] [top]"""
import collections
import csv
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda", "--kernel-name", "regex:" + kern,
                      "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = [i for i, r in enumerate(rows) if "Warp Stall Sampling (All Samples)" in r]
if not hdr:
    print(out[:2000])
    sys.exit(1)
H = rows[hdr[0]]
isamp = H.index("Warp Stall Sampling (All Samples)")
iex = H.index("Instructions Executed") if "Instructions Executed" in H else None
isrc = H.index("Source")
# the view with file/line columns
iloc = None
for name in ("File Path", "File Name", "Source File"):
    if name in H:
        iloc = H.index(name)
iline = H.index("Line") if "Line" in H else None
agg, ex = collections.Counter(), collections.Counter()
tot = 0
for r in rows[hdr[0] + 1:]:
    if len(r) <= isamp or not r[isamp].replace(",", "").isdigit():
        continue
    s = int(r[isamp].replace(",", ""))
    key = (r[iloc].split("/")[-1] + ":" + r[iline]) if (iloc is not None and iline is not None) else r[isrc][:60]
    agg[key] += s
    tot += s
    if iex is not None and r[iex].replace(",", "").isdigit():
        ex[key] += int(r[iex].replace(",", ""))
print(f"# {rep} {kern}: warp-stall samples by source line (total {tot})")
for k, v in agg.most_common(top):
    print(f"{k:50s} samples={v:8d} {v / max(tot, 1):6.3f}  executed={ex[k]}")
