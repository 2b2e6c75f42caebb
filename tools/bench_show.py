#!/usr/bin/env python
"""Pretty-print a bench.py JSON line.  usage: tools/bench_show.py gpurun_out/bench_full.json"""
import json, sys
d = json.load(open(sys.argv[1]))
print("ms/step %.4f  value %.3e  serial %.4f" % (d["ms_per_step"], d["value"], d["aux"]["ms_per_step_branches_serialised"]))
print("roofline", d["roofline"]["kernel"], round(d["roofline"]["frac"], 4), " whole-step frac", d["aux"].get("whole_step_frac_of_peak"))
for n, v in d["aux"]["kernels"].items():
    print("   %-16s %.4f ms  frac %.3f" % (n, v["ms_per_step"], v.get("frac_of_peak", 0)))
print("uncached", d["aux"].get("uncached_plans"))
print("e2e", d.get("e2e"))
t = d["aux"].get("train")
if t:
    print("train:", {k: (round(v, 4) if isinstance(v, float) else v) for k, v in t.items() if not isinstance(v, (dict, str))})
    for n, v in t["kernels"].items():
        print("   %-16s %.4f ms  frac %.3f" % (n, v["ms_per_step"], v.get("frac_of_peak", 0)))
    for k in ("b8_fp32", "b8_fp32_cached_plans"):
        if k in t:
            print(k, round(t[k]["ms_per_step"], 4))
            for n, v in t[k]["kernels"].items():
                print("   %-16s %.4f ms  frac %.3f" % (n, v["ms_per_step"], v.get("frac_of_peak", 0)))
a = d["aux"].get("aten_gpu_baseline")
if a:
    for r in a["runs"]:
        print("aten", {k: (round(v, 3) if isinstance(v, float) and v < 1e6 else v) for k, v in r.items()})
    print("aten peak mem GB", a.get("peak_mem_gb"))
print("cpu", d.get("cpu_baseline"))
print("clocks", d.get("clocks"))
