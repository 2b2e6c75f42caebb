"""GPU: BASELINE.json's full sizes.  Parity against the live torch oracle where it finishes in seconds,
and size-independent properties (batch independence = the multi-GPU sharding invariant, linearity,
determinism) at the bench workload's shape and dtype."""
import pytest
import torch

from helpers import assert_close_scaled
from oracle import torch_path as tp

pytestmark = pytest.mark.gpu
NAMES = ["rgb", "seg", "depth", "bev_rgb", "bev_seg", "bev_height", "voxel_density", "voxel_output"]


def _setup(cfg, B, dtype, mode="val", field="surface", seed=1234):
    from vampire_b200 import ops, synth
    from vampire_b200.matrices import prepare_matrices
    cid = ops.register_config(cfg)
    m = synth.make_mats(cfg, B, mode, seed)
    prep = prepare_matrices(m["sensor2ego_mats"][:, 0], m["intrin_mats"][:, 0], m["ida_mats"][:, 0], m["bda_mat"])
    depth, ctx = synth.make_lift_inputs(cfg, B, seed, dtype)
    vols = synth.make_render_inputs(cfg, B, seed, field=field, dtype=dtype)
    return ops, cid, m, prep, depth, ctx, vols


def test_bench_workload_bf16_vs_live_oracle():
    """configs[1] shape and dtype (R50 256x704, bf16 features), one sample, against the reference's
    PyTorch CPU ops fed the same bf16-rounded inputs."""
    from vampire_b200.config import R50_256x704 as cfg
    ops, cid, m, prep, depth, ctx, (den, sem, feat, rgb) = _setup(cfg, 1, torch.bfloat16)
    conf = cfg.backbone_kwargs()
    buf = tp.build_buffers(conf)
    with torch.no_grad():
        ref_vox = tp.lift_pool(conf, buf, depth.float(), ctx.float(), m)
        ref = tp.render_from_mats(conf, buf, m, den.float(), sem.float(), feat.float(), rgb.float(), torch.tensor(0.1))
    vox, _ = ops.lift_pool_fwd(depth.cuda(), ctx.cuda(), prep.cuda(), cid, True, False, False)
    assert_close_scaled(vox.float().cpu().numpy(), ref_vox.numpy(), 1e-2, "pooled voxel features (bf16)")
    outs = ops.render_fwd(den.cuda(), sem.cuda(), rgb.cuda(), feat.cuda(), torch.tensor(0.1, device="cuda"),
                          prep.cuda(), None, cid, True, 3)
    for n, o, r in zip(NAMES, outs, ref):
        # rendered maps are fp32 from bf16-valued volumes: only voxel_output is rounded to bf16
        assert_close_scaled(o.float().cpu().numpy(), r.numpy(), 1e-2 if n == "voxel_output" else 1e-5, n)


def test_batch_independence_full_size():
    """Sharding invariant: a sample's results do not depend on which batch / rank it rides in.
    B=3 in one call == three B=1 calls, bit for bit, at the bench shape and dtype."""
    from vampire_b200.config import R50_256x704 as cfg
    ops, cid, m, prep, depth, ctx, (den, sem, feat, rgb) = _setup(cfg, 3, torch.bfloat16, mode="train")
    dev = [t.cuda() for t in (depth, ctx, den, sem, feat, rgb)]
    p = prep.cuda()
    beta = torch.tensor(0.1, device="cuda")
    vox, _ = ops.lift_pool_fwd(dev[0], dev[1], p, cid, True, False, False)
    outs = ops.render_fwd(dev[2], dev[3], dev[5], dev[4], beta, p, None, cid, True, 3)
    for b in range(3):
        sl = slice(b, b + 1)
        v1, _ = ops.lift_pool_fwd(dev[0][sl], dev[1][sl], p[sl], cid, True, False, False)
        assert torch.equal(v1[0], vox[b]), f"lift sample {b}"
        o1 = ops.render_fwd(dev[2][sl], dev[3][sl], dev[5][sl], dev[4][sl], beta, p[sl], None, cid, True, 3)
        for n, a, bb in zip(NAMES, o1, outs):
            assert torch.equal(a[0], bb[b]), f"{n} sample {b}"


def test_lift_is_linear_in_ctx_and_depth_full_size():
    """lift+pool is bilinear: scaling ctx by a power of two scales the output exactly (counts unchanged);
    ctx1 + ctx2 gives the sum within fp32 re-association."""
    from vampire_b200.config import R50_256x704 as cfg
    ops, cid, m, prep, depth, ctx, _ = _setup(cfg, 1, torch.float32)
    d, c, p = depth.cuda(), ctx.cuda(), prep.cuda()
    a, _ = ops.lift_pool_fwd(d, c, p, cid, True, False, False)
    b, _ = ops.lift_pool_fwd(d, c * 4.0, p, cid, True, False, False)
    assert torch.equal(b, a * 4.0)
    c2 = torch.randn_like(c)
    s, _ = ops.lift_pool_fwd(d, c + c2, p, cid, True, False, False)
    a2, _ = ops.lift_pool_fwd(d, c2, p, cid, True, False, False)
    assert_close_scaled(s.cpu().numpy(), (a + a2).cpu().numpy(), 1e-5, "additivity in ctx")


def test_render_opaque_and_empty_limits_full_size():
    """Closed-form limits: an empty field (sigma ~ 1e-8) renders the background depth d_bound[1] up to the
    masked-sample density; a solid field renders the first mid-depth and semantics = the first sample."""
    from vampire_b200.config import R50_256x704 as cfg
    ops, cid, m, prep, depth, ctx, (den, sem, feat, rgb) = _setup(cfg, 1, torch.float32, field="empty")
    beta = torch.tensor(0.1, device="cuda")
    p = prep.cuda()
    outs = ops.render_fwd(den.cuda(), sem.cuda(), rgb.cuda(), feat.cuda(), beta, p, None, cid, True, 1)
    depth_map = outs[2]
    assert depth_map.max().item() <= cfg.d_bound[1] + 1e-3
    assert depth_map.min().item() > cfg.d_bound[1] - 1.5      # only the sigma(0) ~ 2.3e-4 haze of masked samples
    solid = torch.full_like(den, -11.0)                        # the reference's init bias (BV2:241): sigma ~ 10
    outs = ops.render_fwd(solid.cuda(), sem.cuda(), rgb.cuda(), feat.cuda(), beta, p, None, cid, True, 1)
    mid0 = 0.5 * (cfg.d_bound[0] + cfg.d_bound[0] + cfg.d_bound[2])
    assert abs(outs[2].median().item() - mid0) < 0.05


@pytest.mark.parametrize("part", ["lift", "render"])
def test_scaled_frustum_config_vs_live_oracle(part):
    """configs[3]: 6 x 512x1408 input (128x352 feature map), same voxel grid, fp32."""
    from vampire_b200.config import R50_512x1408 as cfg
    ops, cid, m, prep, depth, ctx, (den, sem, feat, rgb) = _setup(cfg, 1, torch.float32)
    conf = cfg.backbone_kwargs()
    buf = tp.build_buffers(conf)
    if part == "lift":
        with torch.no_grad():
            ref = tp.lift_pool(conf, buf, depth, ctx, m)
        vox, _ = ops.lift_pool_fwd(depth.cuda(), ctx.cuda(), prep.cuda(), cid, True, False, False)
        assert_close_scaled(vox.cpu().numpy(), ref.numpy(), 1e-5, "vox 512x1408")
    else:
        with torch.no_grad():
            ref = tp.render_from_mats(conf, buf, m, den, sem, feat, rgb, torch.tensor(0.1))
        outs = ops.render_fwd(den.cuda(), sem.cuda(), rgb.cuda(), feat.cuda(), torch.tensor(0.1, device="cuda"),
                              prep.cuda(), None, cid, True, 3)
        for n, o, r in zip(NAMES, outs, ref):
            assert_close_scaled(o.cpu().numpy(), r.numpy(), 1e-5, n + " 512x1408")


def test_scaled_frustum_backward_vs_live_oracle():
    """configs[3] backward: 512x1408 input (128x352 feature map), two cameras (the CPU oracle's grid_sampler backward
    is single-threaded per camera), lift + pool gradients against the reference's autograd."""
    from dataclasses import replace
    from vampire_b200.config import R50_512x1408
    from vampire_b200 import synth
    cfg = replace(R50_512x1408, num_cams=2)
    ops, cid, m, prep, depth, ctx, _ = _setup(cfg, 1, torch.float32, mode="train", seed=91)
    conf = cfg.backbone_kwargs()
    buf = tp.build_buffers(conf)
    cot = synth.make_cotangents([(1, cfg.C, cfg.vZ, cfg.vY, cfg.vX)], 91)[0]
    dr, cr = depth.clone().requires_grad_(True), ctx.clone().requires_grad_(True)
    ref = tp.lift_pool(conf, buf, dr, cr, m)
    rd, rc = torch.autograd.grad((ref * cot).sum(), [dr, cr])
    d, c = depth.cuda().requires_grad_(True), ctx.cuda().requires_grad_(True)
    vox, _ = ops.lift_pool_fwd(d, c, prep.cuda(), cid, True, False, True)
    gd, gc = torch.autograd.grad((vox * cot.cuda()).sum(), [d, c])
    assert_close_scaled(vox.detach().cpu().numpy(), ref.detach().numpy(), 1e-5, "vox 512x1408 (2 cams)")
    assert_close_scaled(gd.cpu().numpy(), rd.numpy(), 2e-5, "d_depth 512x1408")
    assert_close_scaled(gc.cpu().numpy(), rc.numpy(), 2e-5, "d_ctx 512x1408")


@pytest.mark.parametrize("S", [64, 192])
def test_render_sample_sweep_full_size_vs_live_oracle(S):
    """configs[4]: the render at the full R50 ray count with S samples per ray (planes 2.0 + (68 / S) i), camera branch,
    against the reference's ops on the same host."""
    from dataclasses import replace
    from vampire_b200.config import R50_256x704
    cfg = replace(R50_256x704, d_bound=(2.0, 70.0 + 34.0 / S, 68.0 / S))
    assert cfg.S == S
    ops, cid, m, prep, _, _, (den, sem, feat, rgb) = _setup(cfg, 1, torch.float32, seed=55)
    conf = cfg.backbone_kwargs()
    buf = tp.build_buffers(conf)
    with torch.no_grad():
        ref = tp.render_from_mats(conf, buf, m, den, sem, feat, rgb, torch.tensor(0.1))
    beta = torch.tensor(0.1, device="cuda")
    outs = ops.render_fwd(den.cuda(), sem.cuda(), rgb.cuda(), feat.cuda(), beta, prep.cuda(), None, cid, True, 1)
    from vampire_b200.plan import PlanCache
    tab = PlanCache().render(ops.state(cid), cid, prep.cuda(), True).table
    planned = ops.render_fwd(den.cuda(), sem.cuda(), rgb.cuda(), feat.cuda(), beta, prep.cuda(), None, cid, True, 1, tab)
    for n, o, p, r in zip(NAMES[:3], outs, planned, ref):
        assert_close_scaled(o.cpu().numpy(), r.numpy(), 1e-5, f"{n} S={S}")
        assert_close_scaled(p.cpu().numpy(), r.numpy(), 1e-5, f"{n} S={S} (planned march)")
