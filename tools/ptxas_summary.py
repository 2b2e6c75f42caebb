#!/usr/bin/env python
"""Summarise vampire_b200/_lib/obj/*.ptxas.log: kernel, registers, spill bytes, smem.  usage: ptxas_summary.py [filter]"""
import glob, os, re, subprocess, sys
root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "vampire_b200", "_lib", "obj")
flt = sys.argv[1] if len(sys.argv) > 1 else ""
rows = []
for path in sorted(glob.glob(os.path.join(root, "*.ptxas.log"))):
    txt = open(path).read()
    for m in re.finditer(r"Compiling entry function '(\S+)' for 'sm_100a'\n(.*?)\n(.*?)\n(?:.*?Used (\d+) registers(.*))?", txt):
        pass
    blocks = txt.split("Compiling entry function '")[1:]
    for b in blocks:
        name = b.split("'")[0]
        regs = re.search(r"Used (\d+) registers", b)
        spill = re.search(r"(\d+) bytes spill stores, (\d+) bytes spill loads", b)
        smem = re.search(r"(\d+) bytes smem", b)
        rows.append((name, int(regs.group(1)) if regs else -1, spill.group(1) + "/" + spill.group(2) if spill else "?",
                     smem.group(1) if smem else "0"))
names = subprocess.run(["c++filt"], input="\n".join(r[0] for r in rows), capture_output=True, text=True).stdout.splitlines()
for (n, regs, spill, smem), dn in zip(rows, names):
    dn = re.sub(r"\(anonymous namespace\)::", "", dn)
    dn = re.sub(r"\(.*", "", dn).replace("void ", "")
    if flt in dn:
        print(f"{regs:4d} regs  spill {spill:>9}  smem {smem:>6}  {dn}")
