"""GPU: the producer right before the path: softmax over the depth planes (BV2:551),
forward and backward, against torch.softmax (the reference's op) on the CPU."""
import pytest
import torch

from helpers import Case, assert_close_scaled
from oracle import torch_path as tp

pytestmark = pytest.mark.gpu

FP32_REL = 1e-6      # both sides are fp32 exp / sum / divide; only the summation order differs
HALF_REL = 1e-2      # 16-bit outputs


def _mod():
    from vampire_b200.view_transform import LiftRenderB200
    return LiftRenderB200(**Case("mini_val").conf).cuda()


@pytest.mark.parametrize("shape", [(12, 86, 16, 44),      # staged path (inner % 4 == 0)
                                   (2, 6, 17, 8, 12),     # 5-D (B, N, D, fH, fW)
                                   (3, 86, 5, 7),         # inner % 4 != 0 -> scalar path
                                   (2, 1, 4, 4),          # D = 1
                                   (1, 3000, 2, 2)])      # D too large to stage -> scalar path
def test_depth_softmax_fp32(shape):
    mod = _mod()
    g = torch.Generator().manual_seed(11)
    x = torch.randn(*shape, generator=g) * 3.0
    xr = x.clone().requires_grad_(True)
    ref = tp.depth_softmax(xr)
    xc = x.cuda().requires_grad_(True)
    out = mod.depth_softmax(xc)
    assert out.dtype == torch.float32 and out.shape == ref.shape
    assert_close_scaled(out.detach().cpu().numpy(), ref.detach().numpy(), FP32_REL, "softmax")
    assert_close_scaled(out.detach().sum(dim=-3).cpu().numpy(), torch.ones_like(ref.sum(dim=-3)).numpy(), 1e-6, "sums to 1")
    cot = torch.randn(ref.shape, generator=g)
    gr, = torch.autograd.grad((ref * cot).sum(), xr)
    gc, = torch.autograd.grad((out * cot.cuda()).sum(), xc)
    assert_close_scaled(gc.cpu().numpy(), gr.numpy(), 2e-6, "softmax backward")


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("out_fp32", [True, False])
def test_depth_softmax_half_logits(dtype, out_fp32):
    """AMP contract: 16-bit logits, fp32 arithmetic; fp32 probabilities (the reference's autocast) or 16-bit
    probabilities (what the bf16-feature lift consumes).  The gradient returns in the logits' dtype."""
    mod = _mod()
    g = torch.Generator().manual_seed(12)
    x = (torch.randn(6, 86, 16, 44, generator=g) * 2.0).to(dtype)
    xr = x.float().requires_grad_(True)
    ref = tp.depth_softmax(xr)
    xc = x.cuda().requires_grad_(True)
    out = mod.depth_softmax(xc, out_fp32=out_fp32)
    assert out.dtype == (torch.float32 if out_fp32 else dtype)
    assert_close_scaled(out.detach().float().cpu().numpy(), ref.detach().numpy(), FP32_REL if out_fp32 else HALF_REL, "softmax")
    cot = torch.randn(ref.shape, generator=g)
    gr, = torch.autograd.grad((ref * cot).sum(), xr)
    gc, = torch.autograd.grad((out.float() * cot.cuda()).sum(), xc)
    assert gc.dtype == dtype
    assert_close_scaled(gc.float().cpu().numpy(), gr.numpy(), HALF_REL, "softmax backward")


def test_depth_softmax_extreme_logits_and_full_size():
    """Large-magnitude logits must not overflow (max subtraction), and the R50 shape (6 x 86 x 64 x 176) feeds
    the lift: softmax -> lift_pool equals the oracle's softmax -> lift_pool."""
    mod = _mod()
    x = torch.tensor([[-1e4, 0.0, 1e4, 88.0]]).reshape(1, 4, 1, 1).expand(1, 4, 4, 4).contiguous()
    out = mod.depth_softmax(x.cuda())
    assert torch.isfinite(out).all()
    assert_close_scaled(out.cpu().numpy(), tp.depth_softmax(x).numpy(), FP32_REL, "extreme")
    g = torch.Generator().manual_seed(13)
    big = torch.randn(6, 86, 64, 176, generator=g) * 2.0
    out = mod.depth_softmax(big.cuda())
    assert_close_scaled(out.cpu().numpy(), tp.depth_softmax(big).numpy(), FP32_REL, "r50 shape")
