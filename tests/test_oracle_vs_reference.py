"""CPU, build container only: the oracle against the live reference (skipped on the GPU box,
which has no /root/reference -- there the committed golden fixtures pin the oracle)."""
import pytest
import torch

from oracle import torch_path as tp
from oracle.ref_import import build_reference_backbone, reference_available, reference_bilinear_get_voxel_feats
from vampire_b200 import synth
from vampire_b200.config import MINI

pytestmark = pytest.mark.skipif(not reference_available(), reason="/root/reference not present")


@pytest.fixture(scope="module")
def ref():
    return build_reference_backbone(MINI.backbone_kwargs())


@pytest.mark.parametrize("mode", ["val", "train", "stress"])
def test_bit_identical_forward_and_backward(ref, mode):
    cfg, conf = MINI, MINI.backbone_kwargs()
    buf = tp.build_buffers(conf)
    for k in ("frustum", "voxel_coords", "output_coords", "camera_mids", "bev_mids"):
        assert torch.equal(getattr(ref, k), buf[k]), k
    mats = synth.make_mats(cfg, 2, mode, seed=77)
    a = (mats["sensor2ego_mats"][:, 0], mats["intrin_mats"][:, 0], mats["ida_mats"][:, 0], mats["bda_mat"])
    assert torch.equal(ref.get_pixel(*a), tp.get_pixel(buf, *a))
    assert torch.equal(ref.get_geometry(*a), tp.get_geometry(buf, *a))
    depth, ctx = synth.make_lift_inputs(cfg, 2, seed=77)
    den, sem, feat, rgb = synth.make_render_inputs(cfg, 2, seed=77, field="surface")
    leaves = [depth, ctx, den, sem, feat, rgb]
    for t in leaves:
        t.requires_grad_(True)
    geom = torch.nan_to_num(ref.get_geometry(*a), -1e3)
    o_ref = [ref.get_voxel_feats(depth.unsqueeze(2) * ctx.unsqueeze(3), 0, mats)] + \
        list(ref.volume_rendering_from_multiple_views(geom, den, sem, feat, rgb))
    beta = ref.density.beta.detach().clone().requires_grad_(True)
    o_me = [tp.lift_pool(conf, buf, depth, ctx, mats)] + \
        list(tp.volume_rendering(conf, buf, geom, den, sem, feat, rgb, beta))
    for x, y in zip(o_ref, o_me):
        assert torch.equal(x, y)
    cots = synth.make_cotangents([o.shape for o in o_ref], seed=5)
    g_ref = torch.autograd.grad(sum((o * c).sum() for o, c in zip(o_ref, cots)), leaves + [ref.density.beta])
    g_me = torch.autograd.grad(sum((o * c).sum() for o, c in zip(o_me, cots)), leaves + [beta])
    for x, y in zip(g_ref, g_me):
        assert torch.allclose(x, y, rtol=1e-6, atol=1e-7)


def test_no_bda_branch(ref):
    cfg, conf = MINI, MINI.backbone_kwargs()
    buf = tp.build_buffers(conf)
    mats = synth.make_mats(cfg, 1, "val")
    a = (mats["sensor2ego_mats"][:, 0], mats["intrin_mats"][:, 0], mats["ida_mats"][:, 0], None)
    assert torch.equal(ref.get_pixel(*a), tp.get_pixel(buf, *a))
    assert torch.equal(ref.get_geometry(*a), tp.get_geometry(buf, *a))


def test_post_path_restatements(ref):
    """§8f rows 2-3: the Occ3D coordinate buffer and the x4 upsample module equal the reference's own."""
    assert torch.equal(ref.occ_coords, tp.occ_coords())
    x = torch.randn(12, 5, MINI.fH, MINI.fW, generator=torch.Generator().manual_seed(1))
    assert torch.equal(ref.upsample2d(x), tp.upsample(x, MINI.upsample_factor))
    # §8f row 1: the reference's inline `.softmax(dim=1)` on the (B*N, D, fH, fW) depth logits (BV2:551)
    lg = torch.randn(6, MINI.D, MINI.fH, MINI.fW, generator=torch.Generator().manual_seed(2))
    assert torch.equal(lg.softmax(dim=1), tp.depth_softmax(lg))


@pytest.mark.parametrize("mode", ["val", "stress"])
def test_bilinear_2d_lift_restatement(ref, mode):
    """§8f row 4: ``BaseBiLinear.get_voxel_feats`` (base_bilinear.py:471-517), bit for bit incl. its autograd."""
    cfg, conf = MINI, MINI.backbone_kwargs()
    buf = tp.build_buffers(conf)
    mats = synth.make_mats(cfg, 2, mode, seed=31)
    _, ctx = synth.make_lift_inputs(cfg, 2, seed=31)
    a = ctx.clone().requires_grad_(True)
    b = ctx.clone().requires_grad_(True)
    o_ref = reference_bilinear_get_voxel_feats()(ref, a, 0, mats)
    o_me = tp.get_voxel_feats_2d(conf, buf, b, mats)
    assert torch.equal(o_ref, o_me)
    cot = synth.make_cotangents([o_ref.shape], seed=6)[0]
    g_ref, = torch.autograd.grad((o_ref * cot).sum(), a)
    g_me, = torch.autograd.grad((o_me * cot).sum(), b)
    assert torch.allclose(g_ref, g_me, rtol=1e-6, atol=1e-7)


def test_forward_single_sweep_restatement_and_golden():
    """The caller of the path: ``tp.forward_single_sweep`` == the reference's own ``_forward_single_sweep``
    (BV2:518-649) bit for bit, incl. the inline point / occupancy queries (BV2:576-609) that nothing else executes
    from the reference -- and the committed fixture tests/golden/mini_sweep.npz is what that method returns."""
    import numpy as np
    from helpers import load_golden
    from oracle import gen_golden_sweep as gs
    mats, feats, pts = gs.sweep_inputs()
    bb = gs.build_seeded_reference()
    gs.calibrate_density(bb, mats, feats)
    out, rec = gs.run_reference(bb, mats, feats, pts)
    imgs = torch.zeros(gs.BATCH, 1, MINI.num_cams, 3, *MINI.final_dim)
    with torch.no_grad():
        mine = tp.forward_single_sweep(bb, 0, imgs, mats, inrange_pts=pts)
    for a, b in zip(out, mine):
        if isinstance(a, (list, tuple)):
            assert len(a) == len(b) and all(torch.equal(x, y) for x, y in zip(a, b))
        else:
            assert torch.equal(a, b)
    gold = load_golden("mini_sweep")
    outs = gs.flatten_outputs(out)
    for name, a in outs.items():
        st = gs.sample_stride(a.size)
        assert np.array_equal(gold["out_" + name + "_strided"], a.reshape(-1)[::st]), name
    for name in gs.RECORDED:
        assert np.array_equal(gold["rec_" + name + "_out"], rec[name + "_out"].numpy()), name
