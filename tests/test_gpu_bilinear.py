"""GPU: SURVEY §8f row 4 -- the BaseBiLinear ablation's 2-D lift (base_bilinear.py:471-517) as the D = 1 case of
the lift kernels: forward, backward and the bit-exact validity mask, against the torch oracle (which
tests/test_oracle_vs_reference.py pins to the reference's own BaseBiLinear.get_voxel_feats)."""
import pytest
import torch

from helpers import Case, assert_close_scaled
from oracle import torch_path as tp

pytestmark = pytest.mark.gpu


def _mod(case, **kw):
    from vampire_b200.view_transform import LiftRenderB200
    return LiftRenderB200(**case.conf, **kw).cuda()


@pytest.mark.parametrize("name", ["mini_val", "mini_stress"])
def test_lift_2d_forward_backward_fp32(name):
    case = Case(name)
    mod = _mod(case)
    buf = tp.build_buffers(case.conf)
    a = case.ctx.clone().requires_grad_(True)
    ref = tp.get_voxel_feats_2d(case.conf, buf, a, case.mats)
    c = case.ctx.cuda().requires_grad_(True)
    out = mod.lift_pool_2d(c, case.mats)
    assert out.shape == ref.shape
    assert_close_scaled(out.detach().cpu().numpy(), ref.detach().numpy(), 1e-5, "2-D lift")
    cot = torch.randn(ref.shape, generator=torch.Generator().manual_seed(9))
    g_ref, = torch.autograd.grad((ref * cot).sum(), a)
    g_gpu, = torch.autograd.grad((out * cot.cuda()).sum(), c)
    assert_close_scaled(g_gpu.cpu().numpy(), g_ref.numpy(), 2e-5, "2-D lift backward")
    # deterministic (gather) backward
    g2, = torch.autograd.grad((mod.lift_pool_2d(c, case.mats) * cot.cuda()).sum(), c)
    assert torch.equal(g_gpu, g2)


def test_lift_2d_validity_mask_is_bit_exact():
    """valid = (-0.5 < x < W-0.5) & (-0.5 < y < H-0.5) & (z > 0) of the strict projection, for every
    (camera, voxel) pair, and z0 = 0 / iz = 0 everywhere (the single depth plane)."""
    from vampire_b200 import ops
    case = Case("mini_stress")
    cid = ops.register_config(case.cfg, lift_2d=True)
    valid, i0, frac = ops.lift_indices(case.prep.cuda(), cid, True)
    buf = tp.build_buffers(case.conf)
    pix = tp.get_pixel(buf, *case.mat_args())
    H, W = case.cfg.final_dim
    x, y, z = pix[..., 0], pix[..., 1], pix[..., 2]
    exp = (x > -0.5) & (x < W - 0.5) & (y > -0.5) & (y < H - 0.5) & (z > 0.)
    assert torch.equal(valid.cpu().bool(), exp)
    assert int(exp.sum()) > 0
    assert (i0[..., 2].cpu()[exp] == 0).all() and (frac[..., 2].cpu()[exp] == 0).all()


def test_lift_2d_half_features_and_full_size():
    """bf16 features / fp32 accumulation at the R50 geometry (one sample), vs the fp32 oracle on the rounded input."""
    case = Case("r50_val_digest")
    mod = _mod(case)
    buf = tp.build_buffers(case.conf)
    ctx = case.ctx.to(torch.bfloat16)
    ref = tp.get_voxel_feats_2d(case.conf, buf, ctx.float(), case.mats)
    out = mod.lift_pool_2d(ctx.cuda(), case.mats)
    assert out.dtype == torch.bfloat16
    assert_close_scaled(out.float().cpu().numpy(), ref.numpy(), 1e-2, "2-D lift bf16")
    out32 = mod.lift_pool_2d(case.ctx.cuda(), case.mats)
    ref32 = tp.get_voxel_feats_2d(case.conf, buf, case.ctx, case.mats)
    assert_close_scaled(out32.cpu().numpy(), ref32.numpy(), 1e-5, "2-D lift fp32 full size")
