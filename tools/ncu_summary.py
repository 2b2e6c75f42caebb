"""Summarise an .ncu-rep (read here, no GPU needed) into a markdown table + dram traffic JSON.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [more.ncu-rep ...] --out profiles/r01_ncu_summary.md \
        --traffic profiles/dram_traffic.json --dtype bf16
"""
import argparse, csv, io, json, os, re, subprocess, sys

METRICS = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
    ("lts__t_sector_hit_rate.pct", "L2hit%"),
    ("l1tex__t_sector_hit_rate.pct", "L1hit%"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
    ("smsp__inst_executed.sum", "warp_inst"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__waves_per_multiprocessor", "waves"),
]
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1.0}


def rows_of(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    res = []
    for r in rows[2:]:
        d = {"kernel": re.sub(r"void <unnamed>::|\(.*", "", r[idx["Kernel Name"]])}
        for m, short in METRICS:
            if m in idx and r[idx[m]] != "":
                v = float(r[idx[m]].replace(",", ""))
                u = units[idx[m]]
                d[short] = v * UNIT.get(u, 1.0) if short in ("time", "dram_rd", "dram_wr") else v
        res.append(d)
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("reps", nargs="+")
    ap.add_argument("--out", required=True)
    ap.add_argument("--traffic")
    ap.add_argument("--dtype", default="bf16")
    ap.add_argument("--title", default="ncu --set full summary")
    a = ap.parse_args()
    lines = [f"# {a.title}", "",
             "Captured with `ncu --set full --clock-control none --import-source on` under gpurun (1 x B200); "
             "per-launch values; times are cold-cache/serialised replays (compare shares, not absolutes).", "",
             "| kernel | time us | DRAM rd MB | DRAM wr MB | dram % | L2 hit % | L1 hit % | occupancy % | issue % | "
             "warp inst (M) | regs | grid | waves |", "|" + "---|" * 13]
    traffic = {}
    if a.traffic and os.path.exists(a.traffic):
        traffic = json.load(open(a.traffic))
    fresh = set()
    for rep in a.reps:
        for d in rows_of(rep):
            lines.append("| {k} | {t:.1f} | {r:.1f} | {w:.1f} | {dp:.1f} | {l2:.1f} | {l1:.1f} | {oc:.1f} | {iss:.1f} | "
                         "{wi:.2f} | {rg:.0f} | {g:.0f} | {wv:.2f} |".format(
                             k=d["kernel"], t=d.get("time", 0) * 1e6, r=d.get("dram_rd", 0) / 1e6,
                             w=d.get("dram_wr", 0) / 1e6, dp=d.get("dram%", 0), l2=d.get("L2hit%", 0),
                             l1=d.get("L1hit%", 0), oc=d.get("occ%", 0), iss=d.get("issue%", 0),
                             wi=d.get("warp_inst", 0) / 1e6, rg=d.get("regs", 0), g=d.get("grid", 0), wv=d.get("waves", 0)))
            short = {"lift_pool_fwd_kernel": "lift_pool_fwd", "march_fwd_kernel": "march_fwd",
                     "lift_fwd_planned_kernel": "lift_pool_fwd", "march_fwd_planned_kernel": "march_fwd",
                     "pack_cam_volume_kernel": "pack_cam_volume", "bev_channels_vec4_kernel": "bev_fwd"}
            for pat, name in short.items():
                if d["kernel"].startswith(pat):
                    # several launches may match (the march launches a fast and a NaN-safe variant, one of which
                    # returns at once): keep the one that moved the data
                    t = d.get("dram_rd", 0) + d.get("dram_wr", 0)
                    if not isinstance(traffic.get(name), dict):
                        traffic[name] = {}
                    if a.dtype not in traffic.setdefault(name, {}) or name not in fresh or t > traffic[name][a.dtype]:
                        traffic[name][a.dtype] = t
                    fresh.add(name)
    open(a.out, "w").write("\n".join(lines) + "\n")
    if a.traffic:
        traffic["_source"] = "ncu --set full capture " + ", ".join(os.path.basename(r) for r in a.reps) + \
            " (dram__bytes_read.sum + dram__bytes_write.sum per launch; not live)"
        json.dump(traffic, open(a.traffic, "w"), indent=1)
    print("\n".join(lines))


if __name__ == "__main__":
    main()
