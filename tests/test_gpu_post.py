"""GPU: the callers right after the path: x4 bilinear upsample of the rendered maps
and the point / occupancy queries, forward and backward, against the torch oracle."""
import pytest
import torch

from helpers import Case, assert_close_scaled
from oracle import torch_path as tp

pytestmark = pytest.mark.gpu


def _mod(case):
    from vampire_b200.view_transform import LiftRenderB200
    return LiftRenderB200(**case.conf).cuda()


@pytest.mark.parametrize("shape", [(2, 6, 18, 16, 44), (1, 6, 1, 64, 176), (3, 5, 7)])
def test_upsample_forward_backward(shape):
    case = Case("mini_val")
    mod = _mod(case)
    g = torch.Generator().manual_seed(3)
    x = torch.randn(*shape, generator=g)
    xr = x.clone().requires_grad_(True)
    ref = tp.upsample(xr if x.dim() >= 3 else xr, 4) if x.dim() > 3 else tp.upsample(xr[None], 4)[0]
    xc = x.cuda().requires_grad_(True)
    out = mod.upsample2d(xc)
    assert out.shape == ref.shape
    assert_close_scaled(out.detach().cpu().numpy(), ref.detach().numpy(), 1e-6, "upsample")
    cot = torch.randn(ref.shape, generator=g)
    gr, = torch.autograd.grad((ref * cot).sum(), xr)
    gc, = torch.autograd.grad((out * cot.cuda()).sum(), xc)
    assert_close_scaled(gc.cpu().numpy(), gr.numpy(), 1e-5, "upsample backward")
    # deterministic gather backward
    gc2, = torch.autograd.grad((mod.upsample2d(xc) * cot.cuda()).sum(), xc)
    assert torch.equal(gc, gc2)


def test_point_queries_vs_oracle():
    """LiDAR-point queries incl. points outside the volume (border padding / zero padding * valid)."""
    case = Case("mini_stress")
    mod = _mod(case)
    g = torch.Generator().manual_seed(5)
    P = 5000
    pts = torch.rand(P, 3, generator=g) * torch.tensor([120.0, 120.0, 12.0]) - torch.tensor([60.0, 60.0, 7.0])
    sem = case.sem.clone().requires_grad_(True)
    den = case.den.clone().requires_grad_(True)
    cl = torch.randn(case.batch, P, case.cfg.K, generator=g)
    cs = torch.randn(case.batch, P, generator=g)
    ref_l, ref_s = zip(*[tp.point_queries(case.conf, sem, den, pts, i) for i in range(case.batch)])
    ref_l, ref_s = torch.stack(ref_l), torch.stack(ref_s)
    rg = torch.autograd.grad((ref_l * cl).sum() + (ref_s * cs).sum(), [sem, den])
    semc = case.sem.cuda().requires_grad_(True)
    denc = case.den.cuda().requires_grad_(True)
    got_l, got_s = mod.query_points(semc, denc, pts.cuda())
    assert_close_scaled(got_l.detach().cpu().numpy(), ref_l.detach().numpy(), 1e-5, "pts_logits")
    assert_close_scaled(got_s.detach().cpu().numpy(), ref_s.detach().numpy(), 1e-5, "pts_sdf")
    gg = torch.autograd.grad((got_l * cl.cuda()).sum() + (got_s * cs.cuda()).sum(), [semc, denc])
    assert_close_scaled(gg[0].cpu().numpy(), rg[0].numpy(), 2e-5, "d semantic_logits")
    assert_close_scaled(gg[1].cpu().numpy(), rg[1].numpy(), 2e-5, "d density_feature")


@pytest.mark.parametrize("name", ["mini_val", "mini_stress"])
def test_occupancy_queries_vs_oracle(name):
    """Occ3D grid (200x200x16) rotated by bda, semantic logits with border padding and tanh(sigma)."""
    case = Case(name)
    mod = _mod(case)
    coords = tp.occ_coords()
    bda = case.mats["bda_mat"]
    sem = case.sem.clone().requires_grad_(True)
    den = case.den.clone().requires_grad_(True)
    beta = torch.tensor(0.1, requires_grad=True)
    ref_l, ref_d = tp.occupancy_queries(case.conf, sem, den, bda, beta, coords)
    g = torch.Generator().manual_seed(7)
    cl = torch.randn(ref_l.shape, generator=g)
    cd = torch.randn(ref_d.shape, generator=g)
    rg = torch.autograd.grad((ref_l * cl).sum() + (ref_d * cd).sum(), [sem, den, beta])
    semc = case.sem.cuda().requires_grad_(True)
    denc = case.den.cuda().requires_grad_(True)
    got_l, got_d = mod.occupancy(semc, denc, bda.cuda(), coords.cuda())
    assert got_l.shape == ref_l.shape and got_d.shape == ref_d.shape
    assert_close_scaled(got_l.detach().cpu().numpy(), ref_l.detach().numpy(), 1e-5, "occ_logits")
    assert_close_scaled(got_d.detach().cpu().numpy(), ref_d.detach().numpy(), 1e-5, "occ_density")
    gg = torch.autograd.grad((got_l * cl.cuda()).sum() + (got_d * cd.cuda()).sum(), [semc, denc, mod.density.beta])
    assert_close_scaled(gg[0].cpu().numpy(), rg[0].numpy(), 2e-5, "d semantic_logits")
    assert_close_scaled(gg[1].cpu().numpy(), rg[1].numpy(), 5e-5, "d density_feature")
    assert abs(gg[2].item() - rg[2].item()) <= 5e-4 * abs(rg[2].item()) + 1e-6


def test_query_valid_mask_bit_exact():
    """The in-range mask of the queries (BV2:587-589) equals the oracle's, bit for bit."""
    from vampire_b200 import ops
    case = Case("mini_val")
    cid = ops.register_config(case.cfg)
    g = torch.Generator().manual_seed(11)
    pts = torch.rand(20000, 3, generator=g) * torch.tensor([104.0, 104.0, 8.4]) - torch.tensor([52.0, 52.0, 5.2])
    # points exactly on the boundary planes
    pts[:6] = torch.tensor([[-51.2, 0, 0], [51.2, 0, 0], [0, -51.2, 0], [0, 51.2, 0], [0, 0, -5.0], [0, 0, 3.0]])
    lo, ext = tp._seg_lo_ext(case.conf)
    n = ((pts - lo) / ext) * 2. - 1.
    ref = ((n >= -1.) & (n <= 1.)).all(-1)
    _, valid = ops.query_points_fwd(case.den.cuda(), pts.cuda(), None, None, cid, False, False, False)
    assert torch.equal(valid[0].cpu().bool(), ref)
