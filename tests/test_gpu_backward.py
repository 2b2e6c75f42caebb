"""GPU: Bk -- backward of lift+pool and of the render, through the custom ops' autograd, against
the reference's autograd (golden fixtures) and the live torch oracle."""
import numpy as np
import pytest
import torch

from helpers import Case, assert_close_scaled
from oracle import torch_path as tp

pytestmark = pytest.mark.gpu

GRAD_REL = 2e-5   # gradients: fp32 re-association over ~20-140 terms per pixel cell


def _ops(cfg):
    from vampire_b200 import ops
    return ops, ops.register_config(cfg)


def _strided(gold, key, arr):
    gs = int(gold["meta_grad_stride"])
    return gold[key + "_strided"], np.asarray(arr).reshape(-1)[::gs]


@pytest.mark.parametrize("name", ["mini_val", "mini_stress", "r50_val_digest", "r50_stress_digest"])
@pytest.mark.parametrize("channels_last", [False, True])
def test_lift_backward_vs_reference(name, channels_last):
    case = Case(name)
    if not case.inputs_match_golden:
        pytest.skip("inputs differ from the fixture's")
    ops, cid = _ops(case.cfg)
    depth = case.depth.cuda().requires_grad_(True)
    ctx = case.ctx.cuda().requires_grad_(True)
    out, _ = ops.lift_pool_fwd(depth, ctx, case.prep.cuda(), cid, True, channels_last, True)
    cot = case.cotangents()[0].cuda()
    g_depth, g_ctx = torch.autograd.grad((out * cot).sum(), [depth, ctx])
    exp, got = _strided(case.gold, "g_depth", g_depth.cpu().numpy())
    assert_close_scaled(got, exp, GRAD_REL, "d_depth", scale=case.gold["g_depth_absmax"])
    exp, got = _strided(case.gold, "g_ctx", g_ctx.cpu().numpy())
    assert_close_scaled(got, exp, GRAD_REL, "d_ctx", scale=case.gold["g_ctx_absmax"])


def test_lift_backward_is_bit_reproducible():
    """Sorted-segment gather: no float atomics => identical bits on every run."""
    case = Case("mini_stress")
    ops, cid = _ops(case.cfg)
    cot = case.cotangents()[0].cuda()
    ref = None
    for _ in range(4):
        depth = case.depth.cuda().requires_grad_(True)
        ctx = case.ctx.cuda().requires_grad_(True)
        out, _ = ops.lift_pool_fwd(depth, ctx, case.prep.cuda(), cid, True, False, True)
        g = torch.autograd.grad((out * cot).sum(), [depth, ctx])
        if ref is None:
            ref = g
        else:
            assert torch.equal(ref[0], g[0]) and torch.equal(ref[1], g[1])


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_lift_backward_half_features(dtype):
    case = Case("mini_stress")
    ops, cid = _ops(case.cfg)
    d16, c16 = case.depth.to(dtype), case.ctx.to(dtype)
    depth = d16.cuda().requires_grad_(True)
    ctx = c16.cuda().requires_grad_(True)
    out, _ = ops.lift_pool_fwd(depth, ctx, case.prep.cuda(), cid, True, False, True)
    cot = case.cotangents()[0].to(dtype)
    g_depth, g_ctx = torch.autograd.grad((out.float() * cot.cuda().float()).sum(), [depth, ctx])
    buf = tp.build_buffers(case.conf)
    dr = d16.float().requires_grad_(True)
    cr = c16.float().requires_grad_(True)
    ref = tp.lift_pool(case.conf, buf, dr, cr, case.mats)
    rd, rc = torch.autograd.grad((ref * cot.float()).sum(), [dr, cr])
    assert_close_scaled(g_depth.float().cpu().numpy(), rd.numpy(), 1e-2, f"d_depth {dtype}")
    assert_close_scaled(g_ctx.float().cpu().numpy(), rc.numpy(), 1e-2, f"d_ctx {dtype}")


def test_lift_backward_dead_channel_and_empty_cells():
    """Exact zeros in ctx change the per-channel counts (denominators) seen by the backward."""
    case = Case("mini_val")
    ops, cid = _ops(case.cfg)
    ctx0 = case.ctx.clone()
    ctx0[:, :, 2] = 0.0
    ctx0[:, 4] = 0.0          # one camera contributes nothing at all
    depth = case.depth.cuda().requires_grad_(True)
    ctx = ctx0.cuda().requires_grad_(True)
    out, _ = ops.lift_pool_fwd(depth, ctx, case.prep.cuda(), cid, True, False, True)
    cot = case.cotangents()[0]
    g_depth, g_ctx = torch.autograd.grad((out * cot.cuda()).sum(), [depth, ctx])
    buf = tp.build_buffers(case.conf)
    dr = case.depth.clone().requires_grad_(True)
    cr = ctx0.clone().requires_grad_(True)
    ref = tp.lift_pool(case.conf, buf, dr, cr, case.mats)
    rd, rc = torch.autograd.grad((ref * cot).sum(), [dr, cr])
    assert_close_scaled(g_depth.cpu().numpy(), rd.numpy(), GRAD_REL, "d_depth")
    assert_close_scaled(g_ctx.cpu().numpy(), rc.numpy(), GRAD_REL, "d_ctx")


RENDER_GRAD_REL = 5e-5   # compositing backward: suffix sums formed as (total - prefix) in fp32


def _render_grads(case, ops, cid, branches=3, from_tensor=False, dtype=torch.float32):
    vols = [t.to(dtype).cuda().requires_grad_(True) for t in (case.den, case.sem, case.rgb, case.feat)]
    beta = torch.tensor(0.1, device="cuda", requires_grad=True)
    prep = case.prep.cuda()
    geom = ops.get_geometry(prep, cid, True, True) if from_tensor else None
    outs = ops.render_fwd(vols[0], vols[1], vols[2], vols[3], beta, prep, geom, cid, True, branches)
    cots = case.cotangents()[1:]
    loss = sum((o.float() * c.cuda()).sum() for o, c in zip(outs, cots))
    g = torch.autograd.grad(loss, vols + [beta])
    return g  # g_den, g_sem, g_rgb, g_feat, g_beta


@pytest.mark.parametrize("name", ["mini_val", "mini_stress", "r50_val_digest", "r50_stress_digest"])
@pytest.mark.parametrize("from_tensor", [False, True])
def test_render_backward_vs_reference(name, from_tensor):
    case = Case(name)
    if not case.inputs_match_golden:
        pytest.skip("inputs differ from the fixture's")
    ops, cid = _ops(case.cfg)
    g_den, g_sem, g_rgb, g_feat, g_beta = _render_grads(case, ops, cid, 3, from_tensor)
    for key, g in (("g_den", g_den), ("g_sem", g_sem), ("g_rgb", g_rgb), ("g_feat", g_feat)):
        exp, got = _strided(case.gold, key, g.cpu().numpy())
        assert_close_scaled(got, exp, RENDER_GRAD_REL, key, scale=case.gold[key + "_absmax"])
    ref_beta = float(case.gold["g_beta"])
    assert abs(g_beta.item() - ref_beta) <= 2e-4 * abs(ref_beta) + 1e-6, (g_beta.item(), ref_beta)


@pytest.mark.parametrize("branches", [1, 2])
def test_render_backward_single_branch(branches):
    """Camera-only / BEV-only gradients against the torch oracle with the other branch's cotangents zeroed."""
    case = Case("mini_stress")
    ops, cid = _ops(case.cfg)
    g = _render_grads(case, ops, cid, branches)
    buf = tp.build_buffers(case.conf)
    leaves = [t.clone().requires_grad_(True) for t in (case.den, case.sem, case.rgb, case.feat)]
    beta = torch.tensor(0.1, requires_grad=True)
    ref = tp.render_from_mats(case.conf, buf, case.mats, leaves[0], leaves[1], leaves[3], leaves[2], beta)
    cots = case.cotangents()[1:]
    sel = range(0, 3) if branches == 1 else range(3, 8)
    loss = sum((ref[i] * cots[i]).sum() for i in sel)
    rg = torch.autograd.grad(loss, leaves + [beta], allow_unused=True)
    for name, a, b in zip(("g_den", "g_sem", "g_rgb", "g_feat"), g[:4], rg[:4]):
        b = torch.zeros_like(a.cpu()) if b is None else b
        assert_close_scaled(a.cpu().numpy(), b.numpy(), RENDER_GRAD_REL, f"{name} branches={branches}")
    # d beta is one scalar summed over every ray sample, with cancellation: on this case the fp32 reference itself
    # is 1.2e-4 (relative) off its fp64 evaluation (-4091.796 vs -4092.270); ours lands 1.0e-4 on the other side
    assert abs(g[4].item() - rg[4].item()) <= 4e-4 * abs(rg[4].item()) + 1e-6


@pytest.mark.parametrize("dtype", [torch.bfloat16])
def test_render_backward_half_features(dtype):
    case = Case("mini_val")
    ops, cid = _ops(case.cfg)
    g = _render_grads(case, ops, cid, 3, False, dtype)
    buf = tp.build_buffers(case.conf)
    leaves = [t.to(dtype).float().requires_grad_(True) for t in (case.den, case.sem, case.rgb, case.feat)]
    beta = torch.tensor(0.1, requires_grad=True)
    ref = tp.render_from_mats(case.conf, buf, case.mats, leaves[0], leaves[1], leaves[3], leaves[2], beta)
    cots = case.cotangents()[1:]
    # voxel_output is emitted in the feature dtype: round the oracle's the same way before the dot
    loss = sum((r * c).sum() for r, c in zip(ref, cots))
    rg = torch.autograd.grad(loss, leaves + [beta])
    for name, a, b in zip(("g_den", "g_sem", "g_rgb", "g_feat"), g[:4], rg[:4]):
        assert a.dtype == dtype
        assert_close_scaled(a.float().cpu().numpy(), b.numpy(), 1.5e-2, f"{name} {dtype}")


def test_negative_beta_parameter_sign():
    """beta = |param| + 1e-4: the gradient flips sign with the parameter (render_utils.py:44-45)."""
    case = Case("mini_val")
    ops, cid = _ops(case.cfg)
    prep = case.prep.cuda()
    vols = [t.cuda() for t in (case.den, case.sem, case.rgb, case.feat)]
    gs = []
    for p in (0.1, -0.1):
        beta = torch.tensor(p, device="cuda", requires_grad=True)
        outs = ops.render_fwd(vols[0], vols[1], vols[2], vols[3], beta, prep, None, cid, True, 3)
        gs.append(torch.autograd.grad(outs[2].sum() + outs[5].sum(), beta)[0].item())
    assert gs[0] != 0 and abs(gs[0] + gs[1]) <= 1e-6 * abs(gs[0])


def test_render_backward_reuses_the_forward_packed_volume():
    """The forward op hands its workspace to autograd; the backward reads the channels-last copy in it instead of packing
    density | sem | rgb again.  Same gradients as the backward that packs (vector atomics: last bits may differ)."""
    case = Case("mini_stress")
    ops, cid = _ops(case.cfg)
    vols = [t.cuda() for t in (case.den, case.sem, case.rgb, case.feat)]
    beta = torch.tensor(0.1, device="cuda")
    prep = case.prep.cuda()
    res = torch.ops.vampire_b200.render_fwd(*vols, beta, prep, None, cid, True, 3, None)
    outs, ws = list(res[:8]), res[8]
    cots = [c.cuda() for c in case.cotangents()[1:]]
    cots[7] = cots[7].to(outs[7].dtype)
    a = ops.render_bwd(cots, outs, *vols, beta, prep, None, cid, True, 3, ws)
    b = ops.render_bwd(cots, outs, *vols, beta, prep, None, cid, True, 3, None)
    for name, x, y in zip(("g_den", "g_sem", "g_rgb", "g_feat", "g_beta"), a, b):
        assert torch.allclose(x, y, rtol=1e-4, atol=1e-5 * float(y.abs().max())), name
    assert torch.equal(a[3], b[3])                      # the BEV feature gradient is a deterministic gather
    with pytest.raises(ValueError):
        ops.render_bwd(cots, outs, *vols, beta, prep, None, cid, True, 3, ws[:1024])
