#!/usr/bin/env python
"""bench.py -- throughput of the Vampire 2D->3D feature path on B200.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's PyTorch CPU path (oracle port)

A *step* is one pass of the hot path over one batch of synthetic nuScenes-shaped input:
lift+pool (depth (x) ctx -> voxel volume) followed by the render (volumes -> 6 camera maps + BEV).
Default workload = BASELINE.json configs[1]: R50 256x704 geometry, forward, batch 8 per GPU,
bf16 features / fp32 accumulation.  `--workload train` = configs[2]: forward+backward, batch 1 per
GPU, fp32, plus the NCCL all-reduce of the flat gradient bucket.

Prints ONE JSON line (rank 0).  `value` = lifted frustum points per second of the whole job with
inputs resident in HBM; `e2e` = the same metric through the public module API from pinned HOST
buffers with the H2D / D2H copies inside the timed region.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="fwd", choices=["fwd", "train"])
    ap.add_argument("--config", default="r50_256x704")
    ap.add_argument("--batch", type=int, default=None, help="samples per GPU (default 8 fwd / 1 train)")
    ap.add_argument("--dtype", default=None, choices=["bf16", "fp32", "fp16"])
    ap.add_argument("--field", default="surface", choices=["surface", "random", "empty"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--channels-last", action="store_true",
                    help="emit the pooled volume in torch.channels_last_3d instead of the reference's NCDHW strides")
    ap.add_argument("--render-group", type=int, default=0, help="samples per pack/march round (0 = all)")
    ap.add_argument("--plans", default="auto", choices=["auto", "on", "off"],
                    help="drive the lift from cached projection/sort plans (auto: on for fwd = validation, whose "
                         "matrices never change; off for train, whose ida changes every step)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# workload description / algorithmic bytes
# ------------------------------------------------------------------------------------------------
def workload_numbers(cfg, batch, esize):
    N, D, fH, fW, C, K = cfg.num_cams, cfg.D, cfg.fH, cfg.fW, cfg.C, cfg.K
    nvox = cfg.vZ * cfg.vY * cfg.vX
    ncol = cfg.oY * cfg.oX
    pts = batch * N * D * fH * fW
    rays = batch * N * fH * fW
    cp = ((K + 4) + 7) // 8 * 8
    bytes_ = {
        # read depth + ctx, write the voxel volume
        "lift_pool_fwd": batch * ((N * D * fH * fW + N * C * fH * fW) * esize + C * nvox * esize),
        # R1 restricted to the consumed channels: read 22 planes, write the padded channels-last copy
        "pack_cam_volume": batch * ((K + 4) * nvox * esize + cp * nvox * esize),
        # read the 22 consumed channels of the volume once, write 22 fp32 maps
        "march_fwd": batch * ((K + 4) * nvox * esize + (K + 4) * N * fH * fW * 4),
        # read 38 channels over the (oZ+1) z-rows touched, write maps + sigma + resampled features
        "bev_fwd": batch * ((K + 4 + C) * (cfg.oZ + 1) * cfg.vY * cfg.vX * esize
                            + (K + 4) * ncol * 4 + cfg.oZ * ncol * 4 + C * cfg.oZ * ncol * esize),
    }
    return pts, rays, bytes_


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 100 ms during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.tmp = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=self.tmp, stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        self.tmp.flush()
        self.tmp.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.tmp.read().strip().splitlines():
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        self.tmp.close()
        os.unlink(self.tmp.name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
# reference arm: the reference's PyTorch CPU path (oracle/torch_path.py restates its ATen calls)
# ------------------------------------------------------------------------------------------------
def cpu_reference_pass(cfg, ncams, workload, seed=1234):
    """Build a closure running one bounded sample (1 batch sample, `ncams` cameras, fp32) of the
    workload on the host cores; returns (closure, frustum points per call, description)."""
    from oracle import torch_path as tp
    from vampire_b200 import synth
    from dataclasses import replace
    sub = replace(cfg, num_cams=ncams)
    conf = sub.backbone_kwargs()
    buf = tp.build_buffers(conf)
    mats = synth.make_mats(sub, 1, "val", seed)
    depth, ctx = synth.make_lift_inputs(sub, 1, seed)
    den, sem, feat, rgb = synth.make_render_inputs(sub, 1, seed, field="surface")
    beta = torch.tensor(0.1)
    train = workload == "train"
    leaves = [depth, ctx, den, sem, feat, rgb, beta]

    def run():
        if train:
            for t in leaves:
                t.requires_grad_(True)
                t.grad = None
        with torch.set_grad_enabled(train):
            vox = tp.lift_pool(conf, buf, depth, ctx, mats)
            rend = tp.render_from_mats(conf, buf, mats, den, sem, feat, rgb, beta)
            if train:
                loss = vox.sum() + sum(r.sum() for r in rend)
                loss.backward()
        return vox

    pts = ncams * sub.D * sub.fH * sub.fW
    desc = (f"1 sample x {ncams}/{cfg.num_cams} cameras of {cfg.final_dim[0]}x{cfg.final_dim[1]}, fp32, "
            f"{'fwd+bwd' if train else 'fwd'} lift+pool+render, torch {torch.__version__} CPU")
    return run, pts, desc


def run_reference_arm(args, cfg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    # size the per-step sample so (steps + warmup) passes end within a few minutes
    probe, _, _ = cpu_reference_pass(cfg, 1, args.workload)
    t0 = time.perf_counter()
    probe()
    t_cam = time.perf_counter() - t0
    budget = 150.0
    ncams = int(max(1, min(cfg.num_cams, budget / max(1e-3, (args.steps + args.warmup) * t_cam))))
    run, pts, desc = cpu_reference_pass(cfg, ncams, args.workload)
    for _ in range(args.warmup):
        run()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        run()
    dt = time.perf_counter() - t0
    value = pts * args.steps / dt
    line = {
        "impl": "reference", "metric": "lifted_frustum_pts_per_s", "value": value, "unit": "frustum pts/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        # the SAME workload description as this repo's arm (the driver compares the two lines); what one timed
        # step actually ran -- a bounded sample of it, fp32 like the reference -- is stated in cpu_baseline.sample
        "config": workload_config(args, cfg, args.batch or (1 if args.workload == "train" else 8),
                                  args.dtype or ("fp32" if args.workload == "train" else "bf16")),
        "cpu_baseline": {"value": value, "unit": "frustum pts/s", "cores": cores, "kind": "port", "sample": desc},
        "e2e": {"value": value, "unit": "frustum pts/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, cfg, batch, dtype):
    return {
        "workload": ("vampire2_r50_256x704 lift+voxel_pool+render " +
                     ("forward" if args.workload == "fwd" else "forward+backward (DP, grad all-reduce)")),
        "geometry": args.config, "cams": cfg.num_cams, "image": list(cfg.final_dim), "depth_planes": cfg.D,
        "voxel_grid": [cfg.vZ, cfg.vY, cfg.vX], "bev_grid": [cfg.oZ, cfg.oY, cfg.oX],
        "context_channels": cfg.C, "classes": cfg.K, "batch_per_gpu": batch, "features": dtype,
        "density_field": args.field, "ida": "val", "l2": "inputs larger than L2 (no flush needed)",
        "pooled_volume_layout": "channels_last_3d" if args.channels_last else "NCDHW (reference strides)",
        "lift_plans": ("cached per distinct matrices (val-mode matrices never change)"
                       if (args.plans == "on" or (args.plans == "auto" and args.workload == "fwd"))
                       else "off: projection + sort recomputed every call"),
    }


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def main():
    args = parse_args()
    from vampire_b200.config import NAMED
    cfg = NAMED[args.config]
    if args.impl == "reference":
        run_reference_arm(args, cfg)
        return

    import torch.distributed as dist
    from vampire_b200 import cabi, ops, synth
    from vampire_b200.view_transform import LiftRenderB200

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # keep stdout for the ONE JSON line: NCCL's own messages (e.g. the version banner under NCCL_DEBUG=VERSION)
        # go to stderr unless the caller already chose a file
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
    train = args.workload == "train"
    batch = args.batch or (1 if train else 8)
    dname = args.dtype or ("fp32" if train else "bf16")
    tdt = {"bf16": torch.bfloat16, "fp32": torch.float32, "fp16": torch.float16}[dname]
    esize = 4 if dname == "fp32" else 2

    mod = LiftRenderB200(**cfg.backbone_kwargs()).to(dev)
    ops.state(mod.cfg_id).render_group = args.render_group
    # per-rank shard of the job: samples rank*batch .. rank*batch+batch-1 (seeded by sample index)
    seed = 1234 + rank * batch
    mats = synth.make_mats(cfg, batch, "val", seed)
    depth_h, ctx_h = synth.make_lift_inputs(cfg, batch, seed, tdt)
    vols_h = synth.make_render_inputs(cfg, batch, seed, field=args.field, dtype=tdt)   # den, sem, feat, rgb
    host_in = [t.pin_memory() for t in (depth_h, ctx_h) + tuple(vols_h)]
    dev_in = [t.to(dev) for t in host_in]
    mats_dev = {k: v.to(dev) for k, v in mats.items()}
    from vampire_b200.matrices import prepare_matrices
    prep = prepare_matrices(mats["sensor2ego_mats"][:, 0], mats["intrin_mats"][:, 0], mats["ida_mats"][:, 0],
                            mats["bda_mat"]).to(dev)
    beta = mod.density.beta
    use_plans = args.plans == "on" or (args.plans == "auto" and not train)
    plan_tab = None
    if use_plans:
        # built once per distinct matrices (here: before the timed region, like the first batch of a val loop)
        plan_tab = mod.plan_cache.lift(ops.state(mod.cfg_id), mod.cfg_id, prep, True).table

    if train:
        for t in dev_in:
            t.requires_grad_(True)
        from vampire_b200.dp import GradBucket, train_step
        bucket = GradBucket(dev, world)
        c = cfg
        shapes = [(batch, c.C, c.vZ, c.vY, c.vX), (batch, c.num_cams, 3, c.fH, c.fW),
                  (batch, c.num_cams, c.K, c.fH, c.fW), (batch, c.num_cams, 1, c.fH, c.fW), (batch, 3, c.oY, c.oX),
                  (batch, c.K, c.oY, c.oX), (batch, 1, c.oY, c.oX), (batch, 1, c.oZ, c.oY, c.oX),
                  (batch, c.C, c.oZ, c.oY, c.oX)]
        cots = [t.to(dev) for t in synth.make_cotangents(shapes, seed)]
        cots[0] = cots[0].to(tdt)
        cots[8] = cots[8].to(tdt)

    def step_device():
        """hot path with inputs resident in HBM (prepared matrices uploaded once, like a val loop
        whose ida/bda never change)"""
        d, c, den, sem, feat, rgb = dev_in
        if not train:
            with torch.no_grad():
                vox, _ = ops.lift_pool_fwd(d, c, prep, mod.cfg_id, True, args.channels_last, False, plan_tab)
                rend = ops.render_fwd(den, sem, rgb, feat, beta, prep, None, mod.cfg_id, True, 3)
            return vox, rend
        return train_step(mod, d, c, (den, sem, feat, rgb), prep, cots, bucket, plan=plan_tab)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    lib = cabi.lib()
    for _ in range(max(3, args.warmup)):
        step_device()
    barrier()

    # ---- timed region: device-resident ---------------------------------------------------------
    sampler = ClockSampler(local) if rank == 0 else None
    launches0 = cabi.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step_device()
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    launches = cabi.launch_count() - launches0
    clocks = sampler.stop() if sampler else None
    t = torch.tensor([ms_total], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = t.item() / args.steps

    # ---- per-kernel device time: the same K steps again with libvb200 recording a CUDA-event pair around
    # every launch on its launching stream, and the BEV/camera branches serialised (the timed region above
    # overlaps them on a side stream, which would smear their individual durations) -------------------
    cabi.render_set_fork(False)
    cabi.trace_enable(True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step_device()
    e1.record()
    barrier()
    ms_serial = e0.elapsed_time(e1) / args.steps
    trace = cabi.trace_collect()
    cabi.trace_enable(False)
    cabi.render_set_fork(True)

    pts, rays, kbytes = workload_numbers(cfg, batch, esize)
    value = world * pts / (ms_step * 1e-3)

    # ---- e2e: public module API from pinned host buffers, H2D + D2H inside the timed region ------
    e2e = None
    if not args.no_e2e and not train:
        with torch.no_grad():
            vox, rend = step_device()
        host_out = [torch.empty(o.shape, dtype=o.dtype).pin_memory() for o in [vox] + list(rend)]
        del vox, rend

        # Three streams, double-buffered: step k's H2D overlaps step k-1's kernels and step k-2's D2H
        # (PCIe is full duplex), the way a prefetching loader feeds the model.  Every step still moves
        # all of its inputs from pinned host memory and all of its outputs back, inside the timed region.
        s_in, s_run, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        nbuf = 2
        dev_bufs = [[torch.empty_like(h, device=dev) for h in host_in] for _ in range(nbuf)]
        ev_in = [torch.cuda.Event() for _ in range(nbuf)]
        ev_run = [torch.cuda.Event() for _ in range(nbuf)]
        ev_out = [torch.cuda.Event() for _ in range(nbuf)]

        def step_e2e(k):
            i = k % nbuf
            with torch.no_grad():
                with torch.cuda.stream(s_in):
                    s_in.wait_event(ev_run[i])            # buffer i's previous consumer has finished
                    for dbuf, h in zip(dev_bufs[i], host_in):
                        dbuf.copy_(h, non_blocking=True)
                    ev_in[i].record(s_in)
                with torch.cuda.stream(s_run):
                    s_run.wait_event(ev_in[i])
                    d, c, den, sem, feat, rgb = dev_bufs[i]
                    vox = mod.lift_pool(d, c, mats)       # host mats_dict: 4x4 prep on the CPU, as the oracle
                    rend = mod.render(mats, den, sem, feat, rgb)
                    ev_run[i].record(s_run)
                outs = [vox] + list(rend)
                with torch.cuda.stream(s_out):
                    s_out.wait_event(ev_run[i])
                    for o, h in zip(outs, host_out):
                        o.record_stream(s_out)
                        h.copy_(o, non_blocking=True)
                    ev_out[i].record(s_out)

        for k in range(2):
            step_e2e(k)
        barrier()
        k = max(4, min(args.steps, 10))
        e0.record()
        for kk in range(k):
            step_e2e(kk)
        for s_ in (s_in, s_run, s_out):
            torch.cuda.current_stream().wait_stream(s_)
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_e2e = t.item() / k
        e2e = {"value": world * pts / (ms_e2e * 1e-3), "unit": "frustum pts/s", "ms_per_step": ms_e2e,
               "h2d_bytes_per_step": int(sum(h.numel() * h.element_size() for h in host_in) + prep.numel() * 4),
               "d2h_bytes_per_step": int(sum(h.numel() * h.element_size() for h in host_out)), "steps": k,
               "pipelining": "3 streams, double-buffered inputs"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (live CUDA-event time from the timed region) --------------
    peak, peak_src = measured_peaks()
    per_kernel = {}
    for name, (ms, cnt) in trace.items():
        calls_per_step = cnt / args.steps
        entry = {"ms_per_step": ms / args.steps, "launches_per_step": calls_per_step}
        if name in kbytes:
            entry["algorithmic_bytes_per_step"] = kbytes[name]
            entry["achieved_gbs"] = kbytes[name] / (ms / args.steps * 1e-3) / 1e9
            entry["frac_of_peak"] = entry["achieved_gbs"] / peak
        per_kernel[name] = entry
    dom = max((n for n in per_kernel if n in kbytes), key=lambda n: per_kernel[n]["ms_per_step"], default=None)
    roofline = None
    if dom:
        k = per_kernel[dom]
        per_launch_bytes = kbytes[dom] / k["launches_per_step"]
        per_launch_s = k["ms_per_step"] * 1e-3 / k["launches_per_step"]
        roofline = {"kernel": dom, "timing": "CUDA events per launch, second pass of the same steps with the render "
                    "branches serialised", "bound": "hbm", "achieved": per_launch_bytes / per_launch_s / 1e9, "peak": peak,
                    "unit": "GB/s", "frac": per_launch_bytes / per_launch_s / 1e9 / peak, "traffic": None,
                    "peak_source": peak_src, "algorithmic_bytes_per_launch": per_launch_bytes,
                    "us_per_launch": per_launch_s * 1e6}
        prof = os.path.join(ROOT, "profiles", "dram_traffic.json")
        if os.path.exists(prof):
            with open(prof) as fh:
                roofline["traffic"] = json.load(fh).get(dom, {}).get(dname)

    line = {
        "metric": "lifted_frustum_pts_per_s", "value": value, "unit": "frustum pts/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32 arithmetic, %s features" % dname, "data": "synthetic",
        "config": workload_config(args, cfg, batch, dname),
        "aux": {"rendered_rays_per_s": world * rays / (ms_step * 1e-3),
                "ray_samples_per_s": world * rays * cfg.S / (ms_step * 1e-3),
                "lift_only_pts_per_s": (pts / (sum(per_kernel[n]["ms_per_step"] for n in ("ctx_to_nhwc", "lift_pool_fwd")
                                                  if n in per_kernel) * 1e-3)) if "lift_pool_fwd" in per_kernel else None,
                "render_only_rays_per_s": (rays / (sum(per_kernel[n]["ms_per_step"] for n in
                                                      ("pack_cam_volume", "march_fwd", "bev_fwd") if n in per_kernel) * 1e-3))
                if "march_fwd" in per_kernel else None,
                "ms_per_step_branches_serialised": ms_serial,
                "kernels": per_kernel},
        "roofline": roofline,
        "gpu_launches": int(launches),
        "clocks": clocks,
        "e2e": e2e,
    }
    if not args.no_cpu_baseline and world == 1:
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        run, cpts, desc = cpu_reference_pass(cfg, cfg.num_cams, args.workload)
        t0 = time.perf_counter()
        run()
        dt = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": cpts / dt, "unit": "frustum pts/s", "cores": cores, "kind": "port",
                                "sample": desc + f"; 1 pass, {dt:.1f} s"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
