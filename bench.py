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
    ap.add_argument("--no-train-probe", action="store_true", help="skip aux.train (configs[2] inside the fwd run)")
    ap.add_argument("--no-aten-baseline", action="store_true", help="skip aux.aten_gpu_baseline")
    ap.add_argument("--no-uncached", action="store_true", help="skip the uncached-plan variant of the workload")
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
        # backward: read d_vox + depth + ctx, write d_depth + d_ctx
        "lift_pool_bwd": batch * (C * nvox * esize + 2 * (N * D * fH * fW + N * C * fH * fW) * esize),
        # read the 22 cotangent maps + the 22 consumed volume channels, write their gradients
        "march_bwd": batch * ((K + 4) * N * fH * fW * 4 + 2 * (K + 4) * nvox * esize),
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
def cpu_model():
    try:
        with open("/proc/cpuinfo") as fh:
            for line in fh:
                if line.startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def cpu_reference_pass(cfg, ncams, workload, seed=1234):
    """Closures running one bounded sample (1 batch sample, `ncams` cameras, fp32) of the workload on the host
    cores with the reference's own ATen CPU calls: returns (run_lift, run_render, frustum points, rays, description)."""
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

    def run_lift():
        leaves = [depth, ctx]
        for t in leaves:
            t.requires_grad_(train)
            t.grad = None
        with torch.set_grad_enabled(train):
            vox = tp.lift_pool(conf, buf, depth, ctx, mats)
            if train:
                vox.sum().backward()
        return vox

    def run_render():
        leaves = [den, sem, feat, rgb, beta]
        for t in leaves:
            t.requires_grad_(train)
            t.grad = None
        with torch.set_grad_enabled(train):
            rend = tp.render_from_mats(conf, buf, mats, den, sem, feat, rgb, beta)
            if train:
                sum(r.sum() for r in rend).backward()
        return rend

    pts = ncams * sub.D * sub.fH * sub.fW
    rays = ncams * sub.fH * sub.fW
    desc = (f"1 sample x {ncams}/{cfg.num_cams} cameras of {cfg.final_dim[0]}x{cfg.final_dim[1]}, fp32, "
            f"{'fwd+bwd' if train else 'fwd'} lift+pool+render, torch {torch.__version__} CPU")
    return run_lift, run_render, pts, rays, desc


def run_reference_arm(args, cfg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    # size the per-step sample so (steps + warmup) passes end within a few minutes
    lift1, rend1, _, _, _ = cpu_reference_pass(cfg, 1, args.workload)
    t0 = time.perf_counter()
    lift1()
    rend1()
    t_cam = time.perf_counter() - t0
    budget = 150.0
    ncams = int(max(1, min(cfg.num_cams, budget / max(1e-3, (args.steps + max(1, args.warmup)) * t_cam))))
    run_lift, run_render, pts, rays, desc = cpu_reference_pass(cfg, ncams, args.workload)
    for _ in range(max(1, args.warmup)):
        run_lift()
        run_render()
    t_lift = t_rend = 0.0
    for _ in range(args.steps):
        t0 = time.perf_counter()
        run_lift()
        t1 = time.perf_counter()
        run_render()
        t2 = time.perf_counter()
        t_lift += t1 - t0
        t_rend += t2 - t1
    dt = t_lift + t_rend
    value = pts * args.steps / dt
    # what this arm actually ran: ONE fp32 sample per step on the host (the reference runs its sampling in fp32
    # also under AMP); the workload keys are this repo's arm's, batch / feature dtype are the true ones
    ran = workload_config(args, cfg, 1, "fp32")
    ran["cams"] = ncams
    ran["plans"] = "n/a (the reference recomputes get_pixel / get_geometry every call)"
    ran["note"] = ("bounded sample of the workload: 1 sample per step instead of the GPU arm's batch; the CPU path "
                   "has no batching benefit (ATen's 3-D grid_sampler parallelises over batch x cameras only), so "
                   "pts/s is per-sample throughput")
    line = {
        "impl": "reference", "metric": "lifted_frustum_pts_per_s", "value": value, "unit": "frustum pts/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": max(1, args.warmup), "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": ran,
        "aux": {"lift_only_pts_per_s": pts * args.steps / t_lift, "render_only_rays_per_s": rays * args.steps / t_rend,
                "ranks_running": 1, "note": "at N > 1 only rank 0 runs this arm: one CPU process, not N"},
        "cpu_baseline": {"value": value, "unit": "frustum pts/s", "cores": cores, "cpu": cpu_model(), "kind": "port",
                         "sample": desc},
        "e2e": {"value": value, "unit": "frustum pts/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit_json(line)


# ------------------------------------------------------------------------------------------------
# same-box GPU comparator: the reference's own ops on the B200 = stock ATen CUDA kernels
# (base_vampire2.py:507 lift grid_sample, :419 camera, :442 BEV, :431-433 cumsum/exp)
# ------------------------------------------------------------------------------------------------
def _best_of(fn, n=3, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    best = float("inf")
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def aten_gpu_baseline(cfg, dev, batches=(1, 8), seed=1234):
    """The oracle's ATen calls (= the reference's, call for call) executed on the B200: lift and render
    separately, fp32 and bf16-valued inputs (up-cast inside the timed region the way autocast does for
    grid_sample), forward and forward+backward, CUDA events, 1 warm-up + best of 3."""
    from oracle import torch_path as tp
    from vampire_b200 import synth
    conf = cfg.backbone_kwargs()
    buf = tp.buffers_to(tp.build_buffers(conf), dev)
    out = {"what": "reference ops (oracle/torch_path.py = base_vampire2.py:553+483-516, 314-349+391-467) on the same "
                   "B200 with stock ATen CUDA kernels; CUDA events, 1 warm-up + best of 3", "runs": []}
    for B in batches:
        try:
            mats = {k: v.to(dev) for k, v in synth.make_mats(cfg, B, "val", seed).items()}
            depth, ctx = (t.to(dev) for t in synth.make_lift_inputs(cfg, B, seed))
            den, sem, feat, rgb = (t.to(dev) for t in synth.make_render_inputs(cfg, B, seed, field="surface"))
            beta = torch.tensor(0.1, device=dev)
            pts = B * cfg.num_cams * cfg.D * cfg.fH * cfg.fW
            rays = B * cfg.num_cams * cfg.fH * cfg.fW
            for feats in ("fp32", "bf16"):
                if feats == "bf16":
                    lin = [t.detach().bfloat16() for t in (depth, ctx)]
                    rin = [t.detach().bfloat16() for t in (den, sem, feat, rgb)]
                else:
                    lin, rin = [depth, ctx], [den, sem, feat, rgb]
                for train in (False, True):
                    def lift():
                        d, c = (t.detach().float().requires_grad_(train) for t in lin)
                        with torch.set_grad_enabled(train):
                            vox = tp.lift_pool(conf, buf, d, c, mats)
                            if train:
                                vox.sum().backward()

                    def render():
                        vs = [t.detach().float().requires_grad_(train) for t in rin]
                        bt = beta.clone().requires_grad_(train)
                        with torch.set_grad_enabled(train):
                            rend = tp.render_from_mats(conf, buf, mats, *vs, bt)
                            if train:
                                sum(r.sum() for r in rend).backward()

                    ms_l = _best_of(lift)
                    ms_r = _best_of(render)
                    out["runs"].append({"batch": B, "features": feats, "pass": "fwd+bwd" if train else "fwd",
                                        "lift_ms": ms_l, "render_ms": ms_r, "step_ms": ms_l + ms_r,
                                        "lift_pts_per_s": pts / (ms_l * 1e-3), "render_rays_per_s": rays / (ms_r * 1e-3),
                                        "step_pts_per_s": pts / ((ms_l + ms_r) * 1e-3)})
            del depth, ctx, den, sem, feat, rgb, lin, rin
        except torch.cuda.OutOfMemoryError:
            out["runs"].append({"batch": B, "error": "CUDA out of memory in the ATen path"})
        torch.cuda.empty_cache()
    out["peak_mem_gb"] = torch.cuda.max_memory_allocated(dev) / 1e9
    return out


def workload_config(args, cfg, batch, dtype):
    return {
        "workload": ("vampire2_r50_256x704 lift+voxel_pool+render " +
                     ("forward" if args.workload == "fwd" else "forward+backward (DP, grad all-reduce)")),
        "geometry": args.config, "cams": cfg.num_cams, "image": list(cfg.final_dim), "depth_planes": cfg.D,
        "voxel_grid": [cfg.vZ, cfg.vY, cfg.vX], "bev_grid": [cfg.oZ, cfg.oY, cfg.oX],
        "context_channels": cfg.C, "classes": cfg.K, "batch_per_gpu": batch, "features": dtype,
        "density_field": args.field, "ida": "val", "l2": "inputs larger than L2 (no flush needed)",
        "pooled_volume_layout": "channels_last_3d" if args.channels_last else "NCDHW (reference strides)",
        "plans": ("lift + camera-march plans cached per distinct matrices (val-mode matrices never change)"
                       if (args.plans == "on" or (args.plans == "auto" and args.workload == "fwd"))
                       else "off: projection + sort recomputed every call"),
    }


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
class Workload:
    """One configuration of the hot path on this rank's GPU: inputs resident in HBM, `step()` = one pass."""

    def __init__(self, args, cfg, dev, world, rank, train, batch, dname, use_plans, allreduce=True, graph=False):
        from vampire_b200 import ops, synth
        from vampire_b200.matrices import prepare_matrices
        from vampire_b200.view_transform import LiftRenderB200
        self.args, self.cfg, self.dev, self.world, self.train, self.batch, self.dname = args, cfg, dev, world, train, batch, dname
        self.ops = ops
        self.tdt = {"bf16": torch.bfloat16, "fp32": torch.float32, "fp16": torch.float16}[dname]
        self.esize = 4 if dname == "fp32" else 2
        self.mod = LiftRenderB200(**cfg.backbone_kwargs()).to(dev)
        self.mod.train(train)
        ops.state(self.mod.cfg_id).render_group = args.render_group
        # per-rank shard of the job: samples rank*batch .. rank*batch+batch-1 (seeded by sample index)
        seed = 1234 + rank * batch
        self.mats = synth.make_mats(cfg, batch, "val", seed)
        depth_h, ctx_h = synth.make_lift_inputs(cfg, batch, seed, self.tdt)
        vols_h = synth.make_render_inputs(cfg, batch, seed, field=args.field, dtype=self.tdt)   # den, sem, feat, rgb
        self.host_in = [t.pin_memory() for t in (depth_h, ctx_h) + tuple(vols_h)]
        self.dev_in = [t.to(dev) for t in self.host_in]
        self.prep = prepare_matrices(self.mats["sensor2ego_mats"][:, 0], self.mats["intrin_mats"][:, 0],
                                     self.mats["ida_mats"][:, 0], self.mats["bda_mat"]).to(dev)
        self.beta = self.mod.density.beta
        self.plan_tab = self.rplan_tab = None
        if use_plans:
            # built once per distinct matrices (before the timed region, like the first batch of a val loop)
            self.plan_tab = self.mod.plan_cache.lift(ops.state(self.mod.cfg_id), self.mod.cfg_id, self.prep, True).table
            if not train:
                self.rplan_tab = self.mod.plan_cache.render(ops.state(self.mod.cfg_id), self.mod.cfg_id, self.prep,
                                                            True).table
        self.allreduce = allreduce
        if train:
            from vampire_b200.dp import GradBucket
            for t in self.dev_in:
                t.requires_grad_(True)
            self.bucket = GradBucket(dev, world if allreduce else 1)
            c = cfg
            shapes = [(batch, c.C, c.vZ, c.vY, c.vX), (batch, c.num_cams, 3, c.fH, c.fW),
                      (batch, c.num_cams, c.K, c.fH, c.fW), (batch, c.num_cams, 1, c.fH, c.fW), (batch, 3, c.oY, c.oX),
                      (batch, c.K, c.oY, c.oX), (batch, 1, c.oY, c.oX), (batch, 1, c.oZ, c.oY, c.oX),
                      (batch, c.C, c.oZ, c.oY, c.oX)]
            self.cots = [t.to(dev) for t in synth.make_cotangents(shapes, seed)]
            self.cots[0] = self.cots[0].to(self.tdt)
            self.cots[8] = self.cots[8].to(self.tdt)
            self.graphed = None
            if graph:
                from vampire_b200.dp import GraphedTrainStep
                d, c, den, sem, feat, rgb = self.dev_in
                self.graphed = GraphedTrainStep(self.mod, d, c, (den, sem, feat, rgb), self.prep, self.cots, self.bucket,
                                                plan=self.plan_tab)

    def step(self):
        """hot path with inputs resident in HBM (prepared matrices uploaded once, like a val loop
        whose ida/bda never change)"""
        d, c, den, sem, feat, rgb = self.dev_in
        ops, mod = self.ops, self.mod
        if not self.train:
            with torch.no_grad():
                vox, _ = ops.lift_pool_fwd(d, c, self.prep, mod.cfg_id, True, self.args.channels_last, False, self.plan_tab)
                rend = ops.render_fwd(den, sem, rgb, feat, self.beta, self.prep, None, mod.cfg_id, True, 3, self.rplan_tab)
            return vox, rend
        if self.graphed is not None:
            return self.graphed()
        from vampire_b200.dp import train_step
        return train_step(mod, d, c, (den, sem, feat, rgb), self.prep, self.cots, self.bucket, plan=self.plan_tab)

    def barrier(self):
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def timed(self, steps, warmup):
        """W warm-up steps, then exactly K steps between barrier + synchronize; CUDA events; max over ranks."""
        for _ in range(warmup):
            self.step()
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.barrier()
        e0.record()
        for _ in range(steps):
            self.step()
        e1.record()
        self.barrier()
        t = torch.tensor([e0.elapsed_time(e1)], device=self.dev, dtype=torch.float64)
        if self.world > 1:
            import torch.distributed as dist
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item() / steps

    def per_kernel(self, steps):
        """The same K steps again with libvb200 recording a CUDA-event pair around every launch on its launching
        stream and the BEV / camera branches serialised (the timed region overlaps them on a side stream, which
        would smear their individual durations).  Returns (ms per step serialised, {family: entry})."""
        from vampire_b200 import cabi
        cabi.render_set_fork(False)
        cabi.trace_enable(True)
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            self.step()
        e1.record()
        self.barrier()
        ms_serial = e0.elapsed_time(e1) / steps
        trace = cabi.trace_collect()
        cabi.trace_enable(False)
        cabi.render_set_fork(True)
        _, _, kbytes = workload_numbers(self.cfg, self.batch, self.esize)
        peak, _ = measured_peaks()
        per_kernel = {}
        for name, (ms, cnt) in trace.items():
            entry = {"ms_per_step": ms / steps, "launches_per_step": cnt / steps}
            if name in kbytes:
                entry["algorithmic_bytes_per_step"] = kbytes[name]
                entry["achieved_gbs"] = kbytes[name] / (ms / steps * 1e-3) / 1e9
                entry["frac_of_peak"] = entry["achieved_gbs"] / peak
            per_kernel[name] = entry
        return ms_serial, per_kernel, kbytes

    def free(self):
        for name in ("graphed", "dev_in", "host_in", "cots", "bucket", "plan_tab", "rplan_tab", "mod", "prep"):
            if hasattr(self, name):
                delattr(self, name)
        torch.cuda.empty_cache()


def train_probe(args, cfg, dev, world, rank, steps, warmup):
    """BASELINE configs[2] measured inside the default run so the driver sees it at every N: forward+backward,
    B = 1 per GPU, fp32, flat-bucket NCCL all-reduce; the all-reduce's exposed time = step with - step without;
    per-kernel backward times.  Plans off (training draws a new ida every step); the cached variant beside it.
    At N = 1 additionally B = 8 fp32, the size at which a backward roofline fraction means something.

    Every rank draws the SAME sample here (the seed of rank 0): the work of this path depends on the data (rays end at
    surfaces, the rig decides how many voxel-camera pairs exist) and one sample per GPU does not average that out --
    on one B200 the samples that ranks 0..7 would draw take 1.66 / 1.85 / 2.08 / 1.74 / 1.88 / 1.70 / 1.86 / 1.73 ms
    (tools/train_seed_variance.py, profiles/r02_train_seed_variance.txt), so a max over ranks of different samples
    measures that spread (2.08 / 1.66), not the scaling.  Weak scaling = identical work per GPU."""
    out = {"what": "configs[2]: lift+pool+render forward+backward, B=1/GPU, fp32, DP with one flat-bucket NCCL "
                   "all-reduce per step; projection/sort plans recomputed every step; every rank runs the sample of "
                   "rank 0's seed (identical work per GPU: the step time of this data-dependent path varies 1.66-2.08 "
                   "ms between samples on one GPU, profiles/r02_train_seed_variance.txt)"}
    rank = 0      # seeds the sample only; the process group, device and bucket are untouched
    pts1 = cfg.num_cams * cfg.D * cfg.fH * cfg.fW
    w = Workload(args, cfg, dev, world, rank, True, 1, "fp32", False, allreduce=True)
    ms_eager = ms = w.timed(steps, warmup)
    _, kern, _ = w.per_kernel(steps)
    w.free()
    launch = "eager: two custom ops + autograd per step (dp.train_step)"
    # the same step replayed from two CUDA graphs around the one NCCL call (dp.GraphedTrainStep): at B = 1 the eager
    # step is launch-bound as soon as several ranks share a host
    try:
        w = Workload(args, cfg, dev, world, rank, True, 1, "fp32", False, allreduce=True, graph=True)
        ms = w.timed(steps, warmup)
        w.free()
        launch = "two CUDA graphs + one NCCL all-reduce call per step (dp.GraphedTrainStep)"
    except Exception as e:      # capture refused (e.g. a backend that cannot be captured): keep the eager number, say why
        out["graph_error"] = f"{type(e).__name__}: {e}"[:300]
    out.update({"batch_per_gpu": 1, "features": "fp32", "ms_per_step": ms, "pts_per_s": world * pts1 / (ms * 1e-3),
                "steps": steps, "warmup": warmup, "launch": launch, "ms_per_step_eager": ms_eager, "kernels": kern})
    try:
        w = Workload(args, cfg, dev, world, rank, True, 1, "fp32", False, allreduce=False, graph="graph_error" not in out)
        ms_no = w.timed(steps, warmup)
        w.free()
    except Exception:
        w = Workload(args, cfg, dev, world, rank, True, 1, "fp32", False, allreduce=False)
        ms_no = w.timed(steps, warmup)
        w.free()
    out["ms_per_step_without_allreduce"] = ms_no
    out["allreduce_exposed_ms"] = ms - ms_no
    w = Workload(args, cfg, dev, world, rank, True, 1, "fp32", True, allreduce=True)
    out["ms_per_step_cached_plans"] = w.timed(steps, warmup)
    w.free()
    if world == 1:
        for plans in (False, True):
            w = Workload(args, cfg, dev, world, rank, True, 8, "fp32", plans, allreduce=True)
            ms8 = w.timed(max(3, steps // 2), warmup)
            _, kern8, _ = w.per_kernel(max(3, steps // 2))
            w.free()
            out["b8_fp32_cached_plans" if plans else "b8_fp32"] = {
                "batch_per_gpu": 8, "features": "fp32", "ms_per_step": ms8, "pts_per_s": 8 * pts1 / (ms8 * 1e-3),
                "kernels": kern8}
    return out


class StdoutGuard:
    """stdout carries exactly ONE line (the JSON): while the bench runs, file descriptor 1 points at stderr, so
    whatever a native library writes there (NCCL's version banner under NCCL_DEBUG=VERSION, ...) cannot precede it."""

    def __init__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)

    def emit(self, line: str):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        print(line, flush=True)
        os.dup2(2, 1)


_GUARD = None


def emit_json(obj):
    line = json.dumps(obj)
    if _GUARD is not None:
        _GUARD.emit(line)
    else:
        print(line, flush=True)


def main():
    global _GUARD
    args = parse_args()
    from vampire_b200.config import NAMED
    cfg = NAMED[args.config]
    _GUARD = StdoutGuard()
    if args.impl == "reference":
        run_reference_arm(args, cfg)
        return

    import torch.distributed as dist
    from vampire_b200 import cabi

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
        if not os.environ.get("NCCL_DEBUG_FILE"):
            os.environ["NCCL_DEBUG_FILE"] = "/dev/stderr"
        dist.init_process_group("nccl", device_id=dev)
    train = args.workload == "train"
    batch = args.batch or (1 if train else 8)
    dname = args.dtype or ("fp32" if train else "bf16")
    use_plans = args.plans == "on" or (args.plans == "auto" and not train)
    warmup = max(3, args.warmup)

    wl = Workload(args, cfg, dev, world, rank, train, batch, dname, use_plans)
    mod, host_in, mats, prep = wl.mod, wl.host_in, wl.mats, wl.prep
    barrier = wl.barrier
    for _ in range(warmup):
        wl.step()
    barrier()

    # ---- timed region: device-resident ---------------------------------------------------------
    # nvidia-smi needs ~0.3 s to deliver its first line and the timed region is ~20 ms: the sampler is started here,
    # kept busy by extra (untimed) warm-up steps until it has produced a sample, runs through the timed region, and is
    # stopped after a short continuation of the same steps
    sampler = ClockSampler(local) if rank == 0 else None
    t_spin = time.perf_counter()
    while time.perf_counter() - t_spin < 0.6:
        wl.step()
    barrier()
    launches0 = cabi.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        wl.step()
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    launches = cabi.launch_count() - launches0
    t_spin = time.perf_counter()
    while time.perf_counter() - t_spin < 0.4:
        wl.step()
    torch.cuda.synchronize()
    clocks = sampler.stop() if sampler else None
    if clocks is not None:
        clocks["window"] = "0.6 s of the same steps before + the timed region + 0.4 s of the same steps after"
    t = torch.tensor([ms_total], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = t.item() / args.steps

    ms_serial, per_kernel, kbytes = wl.per_kernel(args.steps)
    pts, rays, _ = workload_numbers(cfg, batch, wl.esize)
    value = world * pts / (ms_step * 1e-3)

    # the same workload with the projection / sort recomputed on every call (what a loop whose matrices change every
    # step pays), beside the cached-plan headline
    uncached = None
    if use_plans and not args.no_uncached:
        w2 = Workload(args, cfg, dev, world, rank, train, batch, dname, False)
        ms_u = w2.timed(args.steps, warmup)
        _, kern_u, _ = w2.per_kernel(args.steps)
        w2.free()
        uncached = {"ms_per_step": ms_u, "pts_per_s": world * pts / (ms_u * 1e-3),
                    "lift_pool_fwd_ms": kern_u.get("lift_pool_fwd", {}).get("ms_per_step"),
                    "lift_pool_fwd_frac_of_peak": kern_u.get("lift_pool_fwd", {}).get("frac_of_peak")}

    # ---- e2e: public module API from pinned host buffers, H2D + D2H inside the timed region ------
    e2e = None
    if not args.no_e2e and not train:
        with torch.no_grad():
            vox, rend = wl.step()
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
        mod.plans = "eval" if use_plans else "off"      # the module's own plan cache (keyed by the matrices' bytes)

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
               "pipelining": "3 streams, double-buffered inputs",
               "api": "LiftRenderB200.lift_pool / .render with a host mats_dict (4x4 prep + plan-cache lookup per call)"}
        del dev_bufs, host_out
    wl.free()

    # ---- BASELINE configs[2] (train: fwd+bwd, DP all-reduce) inside the same run, at every N ------
    aux_train = None
    if not train and not args.no_train_probe:
        aux_train = train_probe(args, cfg, dev, world, rank, max(5, args.steps // 2), 3)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (live CUDA-event time) -------------------------------------
    peak, peak_src = measured_peaks()
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
                tr = json.load(fh)
            roofline["traffic"] = tr.get(dom, {}).get(dname)
            roofline["traffic_source"] = tr.get("_source", "profiles/dram_traffic.json (ncu --set full capture, not live)")
    step_bytes = sum(kbytes[n] for n in ("lift_pool_fwd", "march_fwd", "bev_fwd")) if not train else None

    line = {
        "metric": "lifted_frustum_pts_per_s", "value": value, "unit": "frustum pts/s", "n_gpus": world,
        "steps": args.steps, "warmup": warmup, "ms_per_step": ms_step, "higher_is_better": True,
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
                "whole_step_algorithmic_gbs": (step_bytes / (ms_step * 1e-3) / 1e9) if step_bytes else None,
                "whole_step_frac_of_peak": (step_bytes / (ms_step * 1e-3) / 1e9 / peak) if step_bytes else None,
                "uncached_plans": uncached,
                "kernels": per_kernel},
        "roofline": roofline,
        "gpu_launches": int(launches),
        "clocks": clocks,
        "e2e": e2e,
    }
    if aux_train is not None:
        line["aux"]["train"] = aux_train
    if not args.no_aten_baseline and world == 1:
        line["aux"]["aten_gpu_baseline"] = aten_gpu_baseline(cfg, dev)
    if not args.no_cpu_baseline and world == 1:
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        run_lift, run_render, cpts, crays, desc = cpu_reference_pass(cfg, cfg.num_cams, args.workload)
        best_l = best_r = float("inf")
        for i in range(4):                      # 1 warm-up + best of 3
            t0 = time.perf_counter()
            run_lift()
            t1 = time.perf_counter()
            run_render()
            t2 = time.perf_counter()
            if i:
                best_l, best_r = min(best_l, t1 - t0), min(best_r, t2 - t1)
        line["cpu_baseline"] = {"value": cpts / (best_l + best_r), "unit": "frustum pts/s", "cores": cores,
                                "cpu": cpu_model(), "kind": "port",
                                "lift_pts_per_s": cpts / best_l, "render_rays_per_s": crays / best_r,
                                "sample": desc + f"; 1 warm-up + best of 3: lift {best_l:.2f} s, render {best_r:.2f} s"}
        if args.workload == "fwd":
            # the backward side of the same CPU path (fwd and fwd+bwd), one un-warmed pass: it is ~10x the
            # forward (grid_sampler_3d_backward) and a best-of-3 would not fit a default run
            bl, br, _, _, _ = cpu_reference_pass(cfg, cfg.num_cams, "train")
            t0 = time.perf_counter()
            bl()
            t1 = time.perf_counter()
            br()
            t2 = time.perf_counter()
            line["cpu_baseline"]["fwd_bwd"] = {
                "lift_s": t1 - t0, "render_s": t2 - t1, "pts_per_s": cpts / (t2 - t0),
                "sample": "same sample, forward+backward with unit cotangents, one un-warmed pass"}
    emit_json(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
