#!/usr/bin/env python
"""Summarise the per-instruction warp-stall samples of one kernel from an .ncu-rep (source page):
   tools/ncu_stalls.py <report.ncu-rep> <kernel regex> [top N]"""
import csv
import subprocess
import sys

rep, pat = sys.argv[1], sys.argv[2]
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + pat],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, data, seen = None, [], 0
for r in rows:
    if r and r[0] == "Kernel Name":
        seen += 1
        continue
    if r and r[0] == "Address":
        hdr = r
        continue
    if seen == 1 and hdr and len(r) == len(hdr):
        data.append(r)
ix = {h: i for i, h in enumerate(hdr)}
tot = sum(int(r[ix["# Samples"]]) for r in data)
print("instructions", len(data), "samples", tot)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {s: sum(int(r[ix[s]]) for r in data) for s in stalls}
for s, v in sorted(agg.items(), key=lambda x: -x[1])[:8]:
    print("  %-24s %7d  %5.1f%%" % (s, v, 100.0 * v / max(tot, 1)))
for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]]))[:topn]:
    why = {s[6:]: int(r[ix[s]]) for s in stalls if int(r[ix[s]]) * 10 > int(r[ix["# Samples"]])}
    print("%6s  %-80s %s" % (r[ix["# Samples"]], r[ix["Source"]].strip()[:80], why))
