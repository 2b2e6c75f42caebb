#!/usr/bin/env python
"""Per CUDA-source-line instruction / stall shares from an .ncu-rep (needs -lineinfo + --import-source on).
    python tools/src_hot.py rep.ncu-rep kernel_regex [top]"""
import csv, io, subprocess, sys
from collections import defaultdict
rep, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name",
                      "regex:" + pat], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = [i for i, r in enumerate(rows) if r and "Instructions Executed" in r]
hdr = rows[hi[0]]
ci = hdr.index("Instructions Executed"); cs = hdr.index("Warp Stall Sampling (All Samples)")
# the page lists one section per source file of the FIRST matching kernel, then the next kernel: stop at the second
# "Kernel Name" row
kn = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
end = kn[1] if len(kn) > 1 else len(rows)
agg = defaultdict(lambda: [0, 0, ""])
cur_line, cur_src, cur_file = None, "", ""
for r in rows[(kn[0] if kn else 0):end]:
    if r and r[0] == "File Name":
        cur_file = r[1].split("/")[-1]
        continue
    if len(r) <= ci or r[0] == "Line No":
        continue
    if r[0]:
        cur_line, cur_src = r[0], r[1]
    try:
        n = int(r[ci]); st = int(r[cs] or 0)
    except ValueError:
        continue
    a = agg[(cur_file, cur_line)]
    a[0] += n; a[1] += st; a[2] = cur_src
tot = sum(a[0] for a in agg.values()); stt = sum(a[1] for a in agg.values()) or 1
print("total warp-inst", tot)
for (f, ln), (n, st, src) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%5.1f%% inst %5.1f%% stall  %s:%s  %s" % (100 * n / tot, 100 * st / stt, f, ln, src.strip()[:100]))
