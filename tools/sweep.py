"""BASELINE.json configs[3] (scaled frustum, fwd+bwd) and configs[4] (render sweep) on one B200.

    python tools/sweep.py > profiles/r02_config_sweep.md
"""
import os, sys
from dataclasses import replace
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vampire_b200 import cabi, ops, synth
from vampire_b200.config import R50_256x704, R50_512x1408
from vampire_b200.matrices import prepare_matrices


def timed(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def setup(cfg, B, dtype, field="surface"):
    cid = ops.register_config(cfg)
    m = synth.make_mats(cfg, B, "val")
    prep = prepare_matrices(m["sensor2ego_mats"][:, 0], m["intrin_mats"][:, 0], m["ida_mats"][:, 0], m["bda_mat"]).cuda()
    depth, ctx = [t.cuda() for t in synth.make_lift_inputs(cfg, B, dtype=dtype)]
    vols = [t.cuda() for t in synth.make_render_inputs(cfg, B, field=field, dtype=dtype)]
    return cid, prep, depth, ctx, vols


print("# Round 2 -- BASELINE.json configs[3] and configs[4] on 1 x B200 (CUDA events, 10 iterations after 3 warm-ups)\n")
print("## configs[3]: scaled frustum 6x512x1408, D=86, grid 20x256x256, B=1, fp32, forward+backward\n")
print("| geometry | lift fwd ms | lift bwd ms | render fwd ms | render bwd ms | frustum pts | lift fwd Gpts/s | render fwd Mrays/s |")
print("|---|---|---|---|---|---|---|---|")
for name, cfg in (("256x704", R50_256x704), ("512x1408", R50_512x1408)):
    cid, prep, depth, ctx, (den, sem, feat, rgb) = setup(cfg, 1, torch.float32)
    beta = torch.tensor(0.1, device="cuda", requires_grad=True)
    t_lf = timed(lambda: ops.lift_pool_fwd(depth, ctx, prep, cid, True, False, True))
    vox, cnt = ops.lift_pool_fwd(depth, ctx, prep, cid, True, False, True)
    g = torch.randn_like(vox)
    t_lb = timed(lambda: ops.lift_pool_bwd(g, depth, ctx, prep, cnt, cid, True))
    t_rf = timed(lambda: ops.render_fwd(den, sem, rgb, feat, beta, prep, None, cid, True, 3))
    outs = ops.render_fwd(den, sem, rgb, feat, beta, prep, None, cid, True, 3)
    gr = [torch.randn_like(o) for o in outs]
    t_rb = timed(lambda: ops.render_bwd(gr, list(outs), den, sem, rgb, feat, beta, prep, None, cid, True, 3, None))
    pts = cfg.num_cams * cfg.D * cfg.fH * cfg.fW
    rays = cfg.num_cams * cfg.fH * cfg.fW
    print(f"| {name} | {t_lf:.3f} | {t_lb:.3f} | {t_rf:.3f} | {t_rb:.3f} | {pts} | {pts / t_lf / 1e6:.2f} | {rays / t_rf / 1e3:.1f} |")

print("\n## configs[4]: standalone render sweep (camera branch only, B=1, bf16 volume, 'surface' field)\n")
print("planes d_i = 2.0 + (68.4/S) i; rays per image = fH x fW of the feature map\n")
print("| rays/image | S=64 | S=85 | S=128 | S=192 | S=256 |   (ms recomputed geometry / ms cached plan ; Mrays/s ; Gsamples/s with the plan)")
print("|---|---|---|---|---|---|")
for fd in ((256, 704), (512, 1408), (1024, 2816)):
    row = []
    for S in (64, 85, 128, 192, 256):
        step = 68.4 / S
        cfg = replace(R50_256x704, final_dim=fd, d_bound=(2.0, 70.4, step))
        if cfg.S != S:   # arange end-point rounding: nudge the step
            cfg = replace(cfg, d_bound=(2.0, 70.4 + (S - cfg.S) * step * 0.5, step))
        cid, prep, depth, ctx, (den, sem, feat, rgb) = setup(cfg, 1, torch.bfloat16)
        beta = torch.tensor(0.1, device="cuda")
        t = timed(lambda: ops.render_fwd(den, sem, rgb, feat, beta, prep, None, cid, True, 1))
        from vampire_b200.plan import PlanCache
        tab = PlanCache().render(ops.state(cid), cid, prep, True).table
        tp_ = timed(lambda: ops.render_fwd(den, sem, rgb, feat, beta, prep, None, cid, True, 1, tab))
        rays = cfg.num_cams * cfg.fH * cfg.fW
        row.append(f"{t:.3f} / {tp_:.3f} ; {rays / tp_ / 1e3:.0f} ; {rays * cfg.S / tp_ / 1e6:.2f} (S={cfg.S})")
        del tab
        del depth, ctx, den, sem, feat, rgb
    print(f"| {fd[0] // 4}x{fd[1] // 4} | " + " | ".join(row) + " |")


print("\n## next rows (SURVEY 8f 2-3): x4 upsample of the rendered maps and Occ3D queries, R50 256x704, B=1, fp32\n")
print("| op | ms | algorithmic MB | GB/s | % of 6551 GB/s |")
print("|---|---|---|---|---|")
from vampire_b200.view_transform import LiftRenderB200
from oracle import torch_path as tp   # coordinates of the Occ3D grid only (checker-side helper)
cfg = R50_256x704
mod = LiftRenderB200(**cfg.backbone_kwargs()).cuda()
maps = torch.randn(1, cfg.num_cams, cfg.cam_channels, cfg.fH, cfg.fW, device="cuda")
t = timed(lambda: mod.upsample2d(maps))
mb = maps.numel() * 4 * (1 + 16) / 1e6
print(f"| upsample2d x4 (22 maps x 6 cams) fwd | {t:.3f} | {mb:.1f} | {mb / t:.0f} | {mb / t / 65.51:.1f} |")
up = mod.upsample2d(maps)
gup = torch.randn_like(up)
t = timed(lambda: ops.upsample_bwd(gup, 4))
print(f"| upsample2d x4 bwd (gather) | {t:.3f} | {mb:.1f} | {mb / t:.0f} | {mb / t / 65.51:.1f} |")
cid, prep, depth, ctx, (den, sem, feat, rgb) = setup(cfg, 1, torch.float32)
coords = tp.occ_coords().cuda()
bda = torch.eye(4, device="cuda")[None]
t = timed(lambda: mod.occupancy(sem, den, bda, coords))
P = coords.numel() // 3
mb = (P * 3 * 4 + (cfg.K + 1) * P * 4 + (cfg.K + 1) * 16 * 200 * 200 * 4 * 0 + (cfg.K + 1) * cfg.vZ * cfg.vY * cfg.vX * 4 * (16.0 * 0.4 / 8.0) * (80.0 / 102.4) ** 2) / 1e6
print(f"| occupancy queries (640k pts, 18 logits + sigma) fwd: the module call (2 ops + views + tanh) | {t:.3f} | {mb:.1f} | {mb / t:.0f} | {mb / t / 65.51:.1f} |")
# the two query kernels alone (libvb200's per-launch CUDA events), with the grid queried x-fastest (what the module
# does) and in the z-fastest order it arrives in
for label, pts in (("x-fastest (module)", coords.permute(2, 1, 0, 3).reshape(-1, 3).contiguous()),
                   ("z-fastest (as given)", coords.reshape(-1, 3).contiguous())):
    rot = bda[:, :3, :3].contiguous()
    def both():
        ops.query_points_fwd(sem, pts, rot, None, mod.cfg_id, True, False, False)
        ops.query_points_fwd(den, pts, rot, mod.density.beta, mod.cfg_id, False, True, False)
    for _ in range(3):
        both()
    torch.cuda.synchronize()
    cabi.trace_enable(True)
    for _ in range(10):
        both()
    torch.cuda.synchronize()
    tr = cabi.trace_collect()
    cabi.trace_enable(False)
    tk = sum(ms for ms, _ in tr.values()) / 10
    print(f"| ... query kernels only, points {label} | {tk:.3f} | {mb:.1f} | {mb / tk:.0f} | {mb / tk / 65.51:.1f} |")

print("\n## next row (SURVEY 8f 1): softmax over the 86 depth planes (BV2:551), R50 256x704\n")
print("| case | ms | algorithmic MB | GB/s | % of 6551 GB/s |")
print("|---|---|---|---|---|")
for B, dt, out_fp32 in ((1, torch.float32, True), (8, torch.float32, True), (8, torch.bfloat16, False), (8, torch.float16, True)):
    lg = (torch.randn(B * cfg.num_cams, cfg.D, cfg.fH, cfg.fW, device="cuda") * 2).to(dt)
    t = timed(lambda: ops.depth_softmax_fwd(lg, out_fp32))
    es_in, es_out = lg.element_size(), (4 if out_fp32 else lg.element_size())
    mb = lg.numel() * (es_in + es_out) / 1e6
    print(f"| fwd B={B} {str(dt)[6:]} -> {'float32' if out_fp32 else str(dt)[6:]} | {t:.3f} | {mb:.1f} | {mb / t:.0f} | {mb / t / 65.51:.1f} |")
    y = ops.depth_softmax_fwd(lg, out_fp32)
    gy = torch.randn_like(y)
    t = timed(lambda: ops.depth_softmax_bwd(y, gy, lg.dtype))
    mb = (2 * y.numel() * y.element_size() + lg.numel() * es_in) / 1e6
    print(f"| bwd B={B} | {t:.3f} | {mb:.1f} | {mb / t:.0f} | {mb / t / 65.51:.1f} |")
    # same-box comparator: ATen's softmax kernel on the B200
    t = timed(lambda: torch.softmax(lg.float() if out_fp32 else lg, dim=1))
    print(f"| (ATen softmax on the same B200, B={B} {str(dt)[6:]}) | {t:.3f} | | | |")
    del lg, y, gy
