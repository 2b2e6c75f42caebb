"""GPU: the reference's ablation modes of the render -- density_mode='naive' (``self.density = nn.Sigmoid()``,
BV2:191-192, the constructors' default) and cat_seg=True (BV2:449-450, the default of BaseLSSImpaintor) -- against the
reference's own outputs and autograd gradients (tests/golden/mini_naive_catseg.npz, oracle/gen_golden_modes.py)."""
import dataclasses

import numpy as np
import pytest
import torch
import torch.nn as nn

from helpers import assert_close_scaled, golden_value, load_golden
from oracle import gen_golden_modes as gm
from oracle import torch_path as tp
from vampire_b200 import synth
from vampire_b200.config import MINI

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def setup():
    gold = load_golden("mini_naive_catseg")
    mats, den, sem, feat, rgb = gm.inputs()
    chk = np.array([t.double().sum().item() for t in (den, sem, feat, rgb)])
    if not np.allclose(chk, gold["in_checksum"], rtol=1e-7, atol=0):
        pytest.skip("seeded inputs differ from the fixture's on this host")
    return gold, mats, den, sem, feat, rgb


@pytest.fixture()
def fixture_prep(monkeypatch, setup):
    import vampire_b200.view_transform as vt
    prep = torch.from_numpy(setup[0]["prep"])
    monkeypatch.setattr(vt, "prepare_matrices", lambda *a, **k: prep.clone())


def _module(plans="off", **over):
    from vampire_b200.view_transform import LiftRenderB200
    conf = dict(gm.CFG.backbone_kwargs(), **over)
    return LiftRenderB200(plans=plans, **conf).cuda().eval()


@pytest.mark.parametrize("plans", ["off", "always"])
def test_naive_catseg_forward_matches_the_reference(setup, fixture_prep, plans):
    gold, mats, den, sem, feat, rgb = setup
    mod = _module(plans)
    assert isinstance(mod.density, nn.Sigmoid) and not list(mod.parameters())
    with torch.no_grad():
        out = mod.render(mats, den.cuda(), sem.cuda(), feat.cuda(), rgb.cuda())
    assert out[7].shape == (gm.BATCH, gm.CFG.C + gm.CFG.K, gm.CFG.oZ, gm.CFG.oY, gm.CFG.oX)
    for n, o in zip(gm.NAMES, out):
        exp, got = golden_value(gold, "r_" + n, o.float().cpu().numpy())
        assert_close_scaled(got, exp, 1e-5, "naive/cat_seg " + n, scale=float(gold["r_" + n + "_absmax"]))
    # the BEV epilogue (BV2:627-630, 'naive': voxel_output * bev_density) over all 34 channels
    with torch.no_grad():
        epi = mod.render(mats, den.cuda(), sem.cuda(), feat.cuda(), rgb.cuda(), tanh_epilogue=True)
    assert torch.allclose(epi[7], out[7] * out[6], rtol=1e-6, atol=1e-7)
    # ... and fused into the BEV kernel when there is nothing to concatenate
    plain = _module(plans, cat_seg=False)
    with torch.no_grad():
        a = plain.render(mats, den.cuda(), sem.cuda(), feat.cuda(), rgb.cuda())
        b = plain.render(mats, den.cuda(), sem.cuda(), feat.cuda(), rgb.cuda(), tanh_epilogue=True)
    assert torch.equal(a[7], out[7][:, :gm.CFG.C])
    assert torch.allclose(b[7], a[7] * a[6], rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("dtype,rel", [(torch.float32, 1e-5), (torch.bfloat16, 1e-2)])
def test_naive_catseg_backward_matches_the_reference(setup, fixture_prep, dtype, rel):
    gold, mats, den, sem, feat, rgb = setup
    mod = _module("off").train()
    leaves = [t.detach().to(dtype).cuda().requires_grad_(True) for t in (den, sem, feat, rgb)]
    out = mod.render(mats, leaves[0], leaves[1], leaves[2], leaves[3])
    cots = synth.make_cotangents([(1,)] + [tuple(o.shape) for o in out])[1:]
    loss = sum((o.float() * c.cuda()).sum() for o, c in zip(out, cots))
    grads = torch.autograd.grad(loss, leaves)
    if dtype == torch.float32:
        for n, g in zip(("g_den", "g_sem", "g_feat", "g_rgb"), grads):
            exp, got = golden_value(gold, n, g.float().cpu().numpy())
            assert_close_scaled(got, exp, rel, "naive/cat_seg " + n, scale=float(gold[n + "_absmax"]))
    else:
        # bf16-valued inputs: the oracle in fp32 on the same rounded values
        conf = gm.CFG.backbone_kwargs()
        buf = tp.build_buffers(conf)
        ref_leaves = [t.detach().float().cpu().requires_grad_(True) for t in leaves]
        ref = tp.render_from_mats(conf, buf, mats, *ref_leaves, None)
        ref_grads = torch.autograd.grad(sum((r * c).sum() for r, c in zip(ref, cots)), ref_leaves)
        for n, g, r in zip(("g_den", "g_sem", "g_feat", "g_rgb"), grads, ref_grads):
            assert_close_scaled(g.float().cpu().numpy(), r.numpy(), rel, "naive/cat_seg bf16 " + n)


def test_naive_occupancy_matches_the_reference(setup, fixture_prep):
    gold, mats, den, sem, feat, rgb = setup
    mod = _module()
    with torch.no_grad():
        logits, dens = mod.occupancy(sem.cuda(), den.cuda(), mats["bda_mat"], torch.from_numpy(gold["occ_coords"]))
    assert_close_scaled(logits.cpu().numpy(), gold["occ_logits"], 1e-5, "naive occ_logits")
    assert_close_scaled(dens.cpu().numpy(), gold["occ_density_tanh"], 1e-5, "naive occ_density")


def test_sdf_catseg_matches_the_oracle(setup, fixture_prep):
    """cat_seg with the target experiment's density: oracle restatement (bit-pinned above for 'naive')."""
    gold, mats, den, sem, feat, rgb = setup
    cfg = dataclasses.replace(MINI, cat_seg=True)
    from vampire_b200.view_transform import LiftRenderB200
    mod = LiftRenderB200(plans="off", **cfg.backbone_kwargs()).cuda().eval()
    conf = cfg.backbone_kwargs()
    buf = tp.build_buffers(conf)
    with torch.no_grad():
        ref = tp.render_from_mats(conf, buf, mats, den, sem, feat, rgb, torch.tensor(0.1))
        out = mod.render(mats, den.cuda(), sem.cuda(), feat.cuda(), rgb.cuda())
        epi = mod.render(mats, den.cuda(), sem.cuda(), feat.cuda(), rgb.cuda(), tanh_epilogue=True)
    for n, o, r in zip(gm.NAMES, out, ref):
        assert_close_scaled(o.float().cpu().numpy(), r.numpy(), 1e-5, "sdf/cat_seg " + n)
    assert torch.allclose(epi[7], out[7] * out[6].tanh(), rtol=1e-6, atol=1e-7)


class LiveBackbone(nn.Module):
    """Reference attribute names (BV2:127-211) with small live non-path modules, built in the ablation modes."""

    def __init__(self, cfg, img_ch, norm_voxel_coords, feats):
        super().__init__()
        from vampire_b200.view_transform import LiftRenderB200
        for k, v in cfg.backbone_kwargs().items():
            setattr(self, k, v)
        self.cat_pos = True
        self.fD, self.fH, self.fW = cfg.S, cfg.fH, cfg.fW
        lat = LiftRenderB200(**cfg.backbone_kwargs())
        self.register_buffer("camera_mids", lat.camera_mids.clone())
        self.register_buffer("norm_voxel_coords", norm_voxel_coords)
        self.register_buffer("occ_coords", LiftRenderB200.occ_coords())
        self.register_buffer("feats", feats)
        C, K = cfg.C, cfg.K
        self.mapping_along_depth = nn.Conv2d(img_ch, cfg.D, 1)
        self.channel_lower = nn.Conv2d(img_ch, C, 1)
        self.base_conv = nn.Conv3d(C + 3, C, 3, 1, 1)
        self.density_conv = nn.Conv3d(C, 1, 3, 1, 1)
        self.seg_conv = nn.Conv3d(C, K, 3, 1, 1)
        self.rgb_conv = nn.Sequential(nn.Conv3d(C, 3, 3, 1, 1), nn.Sigmoid())
        self.density = nn.Sigmoid()                                               # BV2:191-192
        self.voxel_output = nn.Conv2d((C + K) * cfg.oZ, 24, 1)                    # BV2:199-203 with cat_seg
        self.upsample2d = nn.UpsamplingBilinear2d(scale_factor=cfg.upsample_factor)

    def get_cam_feats(self, imgs):
        return self.feats


class OracleMethods:
    """The path methods of the reference backbone served by the oracle (CPU)."""

    def __init__(self, conf):
        self.conf, self.buf = conf, tp.build_buffers(conf)

    def bind(self, bb):
        bb.get_geometry = lambda s2e, intrin, ida, bda: tp.get_geometry(self.buf, s2e, intrin, ida, bda)
        bb.get_voxel_feats = lambda fr, sweep, mats: tp.get_voxel_feats(self.conf, self.buf, fr, mats)
        bb.volume_rendering_from_multiple_views = lambda geom, den, sem, feat, rgb: tp.volume_rendering(
            self.conf, self.buf, geom, den, sem, feat, rgb, None)
        return bb


@pytest.fixture()
def no_tf32():
    old = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def test_fused_sweep_in_the_ablation_modes(fixture_prep, setup, no_tf32):
    """``fused_forward_single_sweep`` on a backbone built with density_mode='naive', cat_seg=True against the oracle's
    call-for-call restatement of the reference caller (bit-pinned to the reference in test_oracle_vs_reference.py)."""
    from vampire_b200.integration import attach
    from oracle import gen_golden_sweep as gs
    gold, mats = setup[0], setup[1]
    cfg = gm.CFG
    torch.manual_seed(5)
    feats = torch.randn(gm.BATCH, 1, cfg.num_cams, 32, cfg.fH, cfg.fW)
    nvc = torch.from_numpy(load_golden("mini_sweep")["norm_voxel_coords"])
    cpu = LiveBackbone(cfg, 32, nvc, feats).eval()
    g = torch.Generator().manual_seed(11)
    pts = [torch.rand(500, 3, generator=g) * torch.tensor([116.0, 116.0, 10.0]) - torch.tensor([58.0, 58.0, 6.0])
           for _ in range(gm.BATCH)]
    imgs = torch.zeros(gm.BATCH, 1, cfg.num_cams, 3, *cfg.final_dim)
    with torch.no_grad():
        ref = tp.forward_single_sweep(OracleMethods(cfg.backbone_kwargs()).bind(cpu), 0, imgs, mats, inrange_pts=pts)
    gpu = LiveBackbone(cfg, 32, nvc, feats)
    gpu.load_state_dict(cpu.state_dict())
    gpu = attach(gpu.cuda().eval(), fused=True)
    with torch.no_grad():
        out = gpu._forward_single_sweep(0, imgs.cuda(), mats, inrange_pts=[p.cuda() for p in pts])
    assert len(out) == len(ref) == 12 and out[9] == [] and ref[9] == []          # no pts_sdf outside 'sdf' (BV2:593)
    for name, o, r in zip(gs.OUT_NAMES, out, ref):
        if name == "pts_sdf":
            continue
        a = (torch.stack(list(o)) if isinstance(o, (list, tuple)) else o).float().cpu().numpy()
        b = (torch.stack(list(r)) if isinstance(r, (list, tuple)) else r).float().numpy()
        # the live convs run in TF32-free fp32 on both sides, but cuDNN and the CPU kernels sum in different orders
        assert_close_scaled(a, b, 2e-4, "ablation sweep " + name)
