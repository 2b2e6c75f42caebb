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


@pytest.mark.parametrize("name", ["mini_val", "mini_stress", "r50_val_digest"])
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
    assert_close_scaled(got, exp, GRAD_REL, "d_depth")
    exp, got = _strided(case.gold, "g_ctx", g_ctx.cpu().numpy())
    assert_close_scaled(got, exp, GRAD_REL, "d_ctx")


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
