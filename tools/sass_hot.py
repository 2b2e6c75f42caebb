"""Per-opcode and hot-loop SASS statistics from an .ncu-rep source page (needs -lineinfo builds).
    python tools/sass_hot.py rep.ncu-rep kernel_regex [threshold]"""
import csv, io, subprocess, sys
from collections import Counter
rep, pat = sys.argv[1], sys.argv[2]
thr = float(sys.argv[3]) if len(sys.argv) > 3 else 0.6
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + pat],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}; blocks.append(cur)
    elif cur is not None:
        cur["rows"].append(r)
for b in blocks[:1]:
    hdr = b["rows"][0]; idx = {h: i for i, h in enumerate(hdr)}
    data = [r for r in b["rows"][1:] if len(r) > idx["Instructions Executed"]]
    ie = lambda r: int(r[idx["Instructions Executed"]])
    tot = sum(ie(r) for r in data); mx = max(ie(r) for r in data)
    print(b["name"][:110]); print("total warp-inst", tot, "sass lines", len(data))
    op = Counter()
    for r in data:
        s = r[idx["Source"]].split()
        o = s[1] if s[0].startswith("@") else s[0]
        op[o.split(".")[0]] += ie(r)
    print("  ".join(f"{o}:{n / tot * 100:.1f}%" for o, n in op.most_common(18)))
    print(f"--- lines executed >= {thr:.0%} of the hottest ({mx}):")
    for r in data:
        if ie(r) >= thr * mx:
            st = r[idx["Warp Stall Sampling (All Samples)"]]
            print(f"{ie(r) / mx:5.2f} {st:>6} {r[idx['Source']][:110]}")
