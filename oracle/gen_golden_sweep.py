"""ORACLE (test infrastructure only): golden vectors of the REAL ``BaseVAMPIRE2._forward_single_sweep``.

Run in the build container (needs /root/reference):   python -m oracle.gen_golden_sweep

The reference's own ``_forward_single_sweep`` (BV2:518-649) is executed, unmodified, on a real ``BaseVAMPIRE2``
with its real 3-D U-Net and heads (seeded init), MINI geometry, B = 1; only ``get_cam_feats`` is replaced by seeded
features (the image encoder is a stub in this container).  Forward hooks record what the non-path
modules produced (the two lift convs, the U-Net, the three heads, the BEV 1x1 conv) and what the path handed them
(input of ``base_conv``, input of ``voxel_output``).  On the GPU box -- which has no reference -- the tests rebuild a
stand-in backbone whose non-path modules REPLAY these tensors, run the fused drop-in
(``vampire_b200.integration.fused_forward_single_sweep``) and the attached methods on it, and compare the 12-tuple
and the recorded inputs: this pins the point / occupancy queries (BV2:576-609), the x4 upsample (616-626), the BEV
tanh epilogue (627-630) and ``attach()`` against the reference itself, end to end.
-> tests/golden/mini_sweep.npz
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from vampire_b200 import synth  # noqa: E402
from vampire_b200.config import MINI  # noqa: E402
from vampire_b200.matrices import prepare_matrices  # noqa: E402
from oracle.ref_import import build_reference_backbone  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
BATCH, SEED, IMG_CH, NPTS = 1, 9876, 32, 3000
RECORDED = ("mapping_along_depth", "channel_lower", "base_conv", "density_conv", "seg_conv", "rgb_conv", "voxel_output")
OUT_NAMES = ("voxel_output_features", "rgb_preds", "seg_logits_preds", "depth_preds", "bev_rgb_preds",
             "bev_seg_logits_preds", "bev_height_preds", "bev_density", "pts_logits", "pts_sdf", "occ_logits",
             "occ_density")


def sweep_inputs():
    """Seeded inputs shared by the generator and the tests: (mats_dict, image features, in-range points)."""
    mats = synth.make_mats(MINI, BATCH, "stress", seed=SEED)
    g = torch.Generator().manual_seed(SEED)
    feats = torch.randn(BATCH, 1, MINI.num_cams, IMG_CH, MINI.fH, MINI.fW, generator=g)
    # LiDAR-like points: most inside the volume, some outside (border padding / zero padding * valid)
    pts = [torch.rand(NPTS, 3, generator=g) * torch.tensor([116.0, 116.0, 10.0]) - torch.tensor([58.0, 58.0, 6.0])
           for _ in range(BATCH)]
    return mats, feats, pts


def build_seeded_reference():
    torch.manual_seed(SEED)
    conf = MINI.backbone_kwargs()
    bb = build_reference_backbone(dict(conf, img_neck_conf=dict(out_channels=[IMG_CH // 4] * 4)))
    if not hasattr(bb, "voxel_output"):
        # the constructor only builds the BEV 1x1 conv for oY in {128, 256} (BV2:203-209); the MINI grid has oY = 32, so
        # give the unmodified _forward_single_sweep the module it expects (the oY == 128 form)
        bb.voxel_output = torch.nn.Conv2d(MINI.C * MINI.oZ, 80, 1, 1, bias=True)
    return bb.eval()


def calibrate_density(bb, mats, feats):
    """The reference initialises the density bias to sdf_bias - 10 (every ray opaque at the first sample).  Rescale the
    density head (parameters only) so that the SDF feature has mean 0.3 / std 0.8 on these inputs: most of the volume
    is free space and the surface level s = -1 is crossed in places, which makes the rendered maps informative."""
    rec = {}
    h = bb.density_conv.register_forward_hook(lambda m, i, o: rec.__setitem__("d", o.detach()))
    bb.get_cam_feats = lambda imgs: feats
    imgs = torch.zeros(BATCH, 1, MINI.num_cams, 3, *MINI.final_dim)
    with torch.no_grad():
        bb._forward_single_sweep(0, imgs, mats)
        h.remove()
        d = rec["d"]
        scale = 0.8 / d.std().item()
        bb.density_conv.weight.mul_(scale)
        bb.density_conv.bias.copy_((bb.density_conv.bias - d.mean().item()) * scale + 0.3)


def run_reference(bb, mats, feats, pts):
    rec = {}
    hooks = []
    for name in RECORDED:
        def hook(mod, inp, out, name=name):
            rec[name + "_in"] = inp[0].detach().clone()
            rec[name + "_out"] = out.detach().clone()
        hooks.append(getattr(bb, name).register_forward_hook(hook))
    bb.get_cam_feats = lambda imgs: feats
    imgs = torch.zeros(BATCH, 1, MINI.num_cams, 3, *MINI.final_dim)
    with torch.no_grad():
        out = bb._forward_single_sweep(0, imgs, mats, inrange_pts=pts)
    for h in hooks:
        h.remove()
    return out, rec


def sample_stride(size: int, keep: int = 40_000) -> int:
    """Odd stride that keeps about `keep` elements of a flattened array (1 = the full array)."""
    st = max(1, size // keep)
    return st if st % 2 == 1 else st + 1


def flatten_outputs(out):
    d = {}
    for name, o in zip(OUT_NAMES, out):
        d[name] = torch.stack(list(o)).numpy() if isinstance(o, (list, tuple)) else o.numpy()
    return d


def main():
    mats, feats, pts = sweep_inputs()
    bb = build_seeded_reference()
    calibrate_density(bb, mats, feats)
    out, rec = run_reference(bb, mats, feats, pts)
    outs = flatten_outputs(out)
    save = {
        "in_checksum": np.array([feats.double().sum().item(), sum(p.double().sum().item() for p in pts)]),
        "prep": prepare_matrices(mats["sensor2ego_mats"][:, 0], mats["intrin_mats"][:, 0], mats["ida_mats"][:, 0],
                                 mats["bda_mat"]).numpy(),
        "norm_voxel_coords": bb.norm_voxel_coords.numpy(),
        "occ_coords_checksum": np.array(bb.occ_coords.double().sum().item()),
        "beta": np.array(bb.density.beta.item(), dtype=np.float32),
    }
    for name in RECORDED:
        save["rec_" + name + "_out"] = rec[name + "_out"].numpy()
    outs["in_base_conv"] = rec["base_conv_in"].numpy()          # what the path handed the U-Net (lift + cat_pos)
    outs["in_voxel_output"] = rec["voxel_output_in"].numpy()    # ... and the BEV 1x1 conv (tanh epilogue)
    for name, a in outs.items():
        stride = sample_stride(a.size)
        save["out_" + name + "_strided"] = a.reshape(-1)[::stride].copy()
        save["out_" + name + "_max"] = np.array(np.abs(a).max())
    path = os.path.join(GOLDEN, "mini_sweep.npz")
    np.savez_compressed(path, **save)
    print(path, os.path.getsize(path) // 1024, "KiB")
    for name, a in outs.items():
        print(f"  {name:24s} {str(a.shape):28s} max|.| {np.abs(a).max():.4f}")


if __name__ == "__main__":
    main()
