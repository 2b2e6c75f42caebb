"""CPU: the oracle's restatement of the render in the reference's ablation modes (density_mode='naive' = nn.Sigmoid,
cat_seg=True) against tests/golden/mini_naive_catseg.npz -- outputs and autograd gradients of the reference's own
``volume_rendering_from_multiple_views`` and the occupancy query with the density applied (oracle/gen_golden_modes.py)."""
import numpy as np
import torch

from helpers import assert_close_scaled, golden_value, load_golden
from oracle import gen_golden_modes as gm
from oracle import torch_path as tp
from vampire_b200 import synth


def test_oracle_naive_catseg_matches_reference_fixture():
    gold = load_golden("mini_naive_catseg")
    mats, den, sem, feat, rgb = gm.inputs()
    chk = np.array([t.double().sum().item() for t in (den, sem, feat, rgb)])
    assert np.allclose(chk, gold["in_checksum"], rtol=1e-7, atol=0)
    conf = gm.CFG.backbone_kwargs()
    buf = tp.build_buffers(conf)
    for t in (den, sem, feat, rgb):
        t.requires_grad_(True)
    rend = tp.render_from_mats(conf, buf, mats, den, sem, feat, rgb, None)
    assert rend[7].shape[1] == gm.CFG.C + gm.CFG.K                      # cat_seg: 16 feature + 18 logit channels
    for n, r in zip(gm.NAMES, rend):
        exp, got = golden_value(gold, "r_" + n, r.detach().numpy())
        assert_close_scaled(got, exp, 1e-6, "oracle naive/cat_seg " + n, scale=float(gold["r_" + n + "_absmax"]))
    cots = synth.make_cotangents([(1,)] + [r.shape for r in rend])[1:]
    grads = torch.autograd.grad(sum((r * c).sum() for r, c in zip(rend, cots)), [den, sem, feat, rgb])
    for n, g in zip(("g_den", "g_sem", "g_feat", "g_rgb"), grads):
        exp, got = golden_value(gold, n, g.numpy())
        assert_close_scaled(got, exp, 1e-5, "oracle naive/cat_seg " + n, scale=float(gold[n + "_absmax"]))
    with torch.no_grad():
        logits, dens = tp.occupancy_queries(conf, sem, den, mats["bda_mat"], None, torch.from_numpy(gold["occ_coords"]))
    assert_close_scaled(logits.numpy(), gold["occ_logits"], 1e-6, "oracle naive occ_logits")
    assert_close_scaled(dens.numpy(), gold["occ_density_tanh"], 1e-6, "oracle naive occ_density")


def test_det_points_are_the_reference_output_coords_flipped():
    """cat_seg samples the logits at ``LiftRenderB200.det_points()``: they must be the reference's ``output_coords``
    buffer (built by the oracle with the reference's own torch calls) with z flipped, bit for bit, in volume order."""
    from vampire_b200.view_transform import LiftRenderB200
    conf = gm.CFG.backbone_kwargs()
    mod = LiftRenderB200(**conf)
    ref = tp.build_buffers(conf)["output_coords"][..., :3].flip(0).reshape(-1, 3)
    assert torch.equal(mod.det_points(), ref)
    assert isinstance(mod.density, torch.nn.Sigmoid) and mod._beta().numel() == 1
    with torch.no_grad():      # the epilogue of BV2:629-630 in 'naive' mode multiplies by the density itself
        v, d = torch.randn(1, 34, 5, 4, 4), torch.rand(1, 1, 5, 4, 4)
        assert torch.equal(mod.bev_epilogue(v, d), v * d)
