"""ORACLE (test infrastructure only): golden vectors of the reference's ablation modes of the render.

Run in the build container (needs /root/reference):   python -m oracle.gen_golden_modes

``density_mode='naive'`` (``self.density = nn.Sigmoid()``, BV2:191-192 -- the constructors' default) and
``cat_seg=True`` (the resampled semantic logits concatenated behind the BEV features, BV2:449-450 -- the default of
``BaseLSSImpaintor``) are not what the target experiment uses (base_exp.py:51,54), but a backbone built with the
constructor defaults has them.  The reference's own ``volume_rendering_from_multiple_views`` (+ autograd) and the
occupancy query of BV2:597-609 are run on a real ``BaseVAMPIRE2`` built with those modes, MINI geometry.
-> tests/golden/mini_naive_catseg.npz
"""
from __future__ import annotations

import dataclasses
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from vampire_b200 import synth  # noqa: E402
from vampire_b200.config import MINI  # noqa: E402
from vampire_b200.matrices import prepare_matrices  # noqa: E402
from oracle.ref_import import build_reference_backbone  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
CFG = dataclasses.replace(MINI, density_mode="naive", cat_seg=True)
BATCH, MODE, FIELD = 2, "stress", "random"
NAMES = ["rgb", "seg", "depth", "bev_rgb", "bev_seg", "bev_height", "voxel_density", "voxel_output"]
STRIDE = 7        # arrays over 60k elements are stored as every 7th element (+ their max |.|)


def put(out, key, t):
    a = t.detach().numpy()
    if a.size > 60_000:
        out[key + "_strided"] = a.reshape(-1)[::STRIDE].copy()
    else:
        out[key] = a
    out[key + "_absmax"] = float(np.abs(a).max())


def inputs():
    mats = synth.make_mats(CFG, BATCH, MODE)
    den, sem, feat, rgb = synth.make_render_inputs(CFG, BATCH, field=FIELD)
    return mats, den, sem, feat, rgb


def main():
    torch.manual_seed(0)
    bb = build_reference_backbone(CFG.backbone_kwargs())
    assert isinstance(bb.density, torch.nn.Sigmoid) and bb.cat_seg
    mats, den, sem, feat, rgb = inputs()
    args = (mats["sensor2ego_mats"][:, 0], mats["intrin_mats"][:, 0], mats["ida_mats"][:, 0], mats["bda_mat"])
    out = {"in_checksum": np.array([t.double().sum().item() for t in (den, sem, feat, rgb)]),
           "prep": prepare_matrices(*args).numpy(), "meta_stride": STRIDE}
    with torch.no_grad():
        geom = torch.nan_to_num(bb.get_geometry(*args), -1e3)
    for t in (den, sem, feat, rgb):
        t.requires_grad_(True)
    rend = bb.volume_rendering_from_multiple_views(geom, den, sem, feat, rgb)
    for n, r in zip(NAMES, rend):
        put(out, "r_" + n, r)
    cots = synth.make_cotangents([(1,)] + [r.shape for r in rend])[1:]
    loss = sum((r * c).sum() for r, c in zip(rend, cots))
    grads = torch.autograd.grad(loss, [den, sem, feat, rgb])
    for n, g in zip(("g_den", "g_sem", "g_feat", "g_rgb"), grads):
        put(out, n, g)
    # occupancy query with the density applied to the volume first (BV2:597-609, 647-648), on a small grid of points
    with torch.no_grad():
        gen = torch.Generator().manual_seed(77)
        coords = (torch.rand(6, 5, 4, 3, generator=gen) * torch.tensor([110.0, 110.0, 9.0])
                  - torch.tensor([55.0, 55.0, 5.5]))
        lo = torch.as_tensor([bb.x_bound_seg[0], bb.y_bound_seg[0], bb.z_bound_seg[0]])
        ext = torch.as_tensor([bb.x_bound_seg[1] - bb.x_bound_seg[0], bb.y_bound_seg[1] - bb.y_bound_seg[0],
                               bb.z_bound_seg[1] - bb.z_bound_seg[0]])
        rot = mats["bda_mat"][:, :3, :3].view(BATCH, 1, 1, 1, 3, 3)
        c = (rot @ coords[None, ..., None].expand(BATCH, *coords.shape, 1)).squeeze(-1)
        n = ((c - lo) / ext) * 2. - 1.
        occ_logits = F.grid_sample(sem, n, padding_mode='border', align_corners=True)
        occ_density = F.grid_sample(bb.density(den), n, align_corners=True)
        out["occ_coords"] = coords.numpy()
        out["occ_logits"] = occ_logits.permute(0, 2, 3, 4, 1).numpy()
        out["occ_density_tanh"] = occ_density.permute(0, 2, 3, 4, 1).tanh().numpy()
    path = os.path.join(GOLDEN, "mini_naive_catseg.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path) // 1024, "KiB")
    for n, r in zip(NAMES, rend):
        print(f"  {n:14s} {str(tuple(r.shape)):24s} max|.| {out['r_' + n + '_absmax']:.4f}")


if __name__ == "__main__":
    main()
