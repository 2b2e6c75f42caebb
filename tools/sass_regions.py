"""SASS regions by execution count.  python tools/sass_regions.py rep kernel_regex units_per_launch"""
import csv, io, subprocess, sys
rep, pat, units = sys.argv[1], sys.argv[2], float(sys.argv[3])
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + pat], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}; blocks.append(cur)
    elif cur is not None:
        cur["rows"].append(r)
b = [x for x in blocks if pat.split("|")[0] in x["name"]][0]
hdr = b["rows"][0]; idx = {h: i for i, h in enumerate(hdr)}
data = [r for r in b["rows"][1:] if len(r) > idx["Instructions Executed"]]
ie = lambda r: int(r[idx["Instructions Executed"]])
tot = sum(ie(r) for r in data)
print(b["name"][:100]); print("per unit", tot / units, "total", tot)
runs = []
for i, r in enumerate(data):
    n = ie(r) / units
    if runs and abs(runs[-1][2] - n) <= 0.03 * max(n, 0.02): runs[-1][1] = i; runs[-1][3] += 1
    else: runs.append([i, i, n, 1])
for a, bb, n, c in runs:
    if c * n >= float(sys.argv[4]) if len(sys.argv) > 4 else 4:
        print(f"sass[{a:4d}-{bb:4d}] exec/unit {n:6.3f} x {c:4d} = {c * n:7.1f}   first: {data[a][idx['Source']][:70]}")
