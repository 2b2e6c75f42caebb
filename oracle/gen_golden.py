"""ORACLE (test infrastructure only): generate ``tests/golden/*.npz`` from the REAL reference.

Run in the build container (needs /root/reference):   python -m oracle.gen_golden

The reference has no tests or golden vectors of its own, so these fixtures are
outputs of the reference's own methods -- ``get_pixel``, ``get_geometry``, ``get_voxel_feats``,
``volume_rendering_from_multiple_views`` (BV2:314-516) and their autograd -- on the seeded
synthetic inputs of ``vampire_b200.synth``.  Inputs are NOT stored (they are regenerated from the
seed; a checksum guards against RNG drift); bit-exact quantities are stored as SHA-256 digests
plus the arrays needed to localise a mismatch; tolerance-checked quantities as fp32 arrays
(strided for the big ones).
"""
from __future__ import annotations

import hashlib
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from vampire_b200.config import MINI, R50_256x704, PathConfig  # noqa: E402
from vampire_b200.lattice import build_lattice  # noqa: E402
from vampire_b200.matrices import prepare_matrices  # noqa: E402
from vampire_b200 import synth  # noqa: E402
from oracle import strict_np as sn  # noqa: E402
from oracle.ref_import import build_reference_backbone  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def sha(a) -> str:
    a = np.ascontiguousarray(a)
    return hashlib.sha256(a.tobytes()).hexdigest()


def seg_lo_ext(cfg: PathConfig):
    lo = (cfg.x_bound_seg[0], cfg.y_bound_seg[0], cfg.z_bound_seg[0])
    ext = (cfg.x_bound_seg[1] - cfg.x_bound_seg[0], cfg.y_bound_seg[1] - cfg.y_bound_seg[0],
           cfg.z_bound_seg[1] - cfg.z_bound_seg[0])
    return lo, ext


def run_case(name: str, cfg: PathConfig, batch: int, mode: str, field: str, stride_big: int,
             with_backward: bool, store_full_geometry: bool):
    t0 = time.time()
    conf = cfg.backbone_kwargs()
    bb = build_reference_backbone(conf)
    lat = build_lattice(cfg)
    mats = synth.make_mats(cfg, batch, mode)
    depth, ctx = synth.make_lift_inputs(cfg, batch)
    den, sem, feat, rgb = synth.make_render_inputs(cfg, batch, field=field)
    args = (mats["sensor2ego_mats"][:, 0], mats["intrin_mats"][:, 0], mats["ida_mats"][:, 0], mats["bda_mat"])
    out = {
        "meta_name": name, "meta_batch": batch, "meta_mode": mode, "meta_field": field,
        "meta_stride": stride_big, "meta_torch": torch.__version__,
        "in_checksum": np.array([depth.double().sum().item(), ctx.double().sum().item(),
                                 den.double().sum().item(), sem.double().sum().item(),
                                 feat.double().sum().item(), rgb.double().sum().item()]),
        "sensor2ego": args[0].numpy(), "intrin": args[1].numpy(), "ida": args[2].numpy(), "bda": args[3].numpy(),
        "prep": prepare_matrices(*args).numpy(),
    }

    # ---- G1/G2 geometry (bit-exact rows) ---------------------------------------------------
    with torch.no_grad():
        pix = bb.get_pixel(*args).contiguous()
        geom_raw = bb.get_geometry(*args).contiguous()
        geom = torch.nan_to_num(geom_raw, -1e3)
    out["pix_sha"] = sha(pix.numpy())
    out["geom_sha"] = sha(geom_raw.numpy())
    if store_full_geometry:
        out["pix"] = pix.numpy()
        out["geom"] = geom_raw.numpy()
    else:
        gst = max(stride_big, 13)
        out["meta_geom_stride"] = gst
        out["pix_strided"] = pix.numpy().reshape(-1)[::gst].copy()
        out["geom_strided"] = geom_raw.numpy().reshape(-1)[::gst].copy()

    # ---- L2/R2 integer rows, derived from the reference's own coordinates -------------------
    li = sn.lift_indices(pix.numpy(), cfg.final_dim, cfg.d_bound, (cfg.fW, cfg.fH, cfg.D))
    lo, ext = seg_lo_ext(cfg)
    ri = sn.render_indices(geom[:, :, :-1].numpy(), lo, ext, (cfg.vX, cfg.vY, cfg.vZ))
    out["lift_valid_sha"] = sha(li["valid"].astype(np.uint8))
    out["lift_valid_count"] = int(li["valid"].sum())
    l0 = np.stack(li["i0"], -1).astype(np.int16)
    out["lift_i0_sha"] = sha(l0)
    out["lift_i0_valid_sha"] = sha(l0[li["valid"]])
    out["render_mask_sha"] = sha(ri["mask"].astype(np.uint8))
    out["render_mask_count"] = int(ri["mask"].sum())
    r0 = np.stack(ri["i0"], -1)
    out["render_i0_masked_sha"] = sha(r0[ri["mask"]].astype(np.int16))

    # ---- L1-L4 lift + pool, R1-R6 render, Bk backward ----------------------------------------
    depth.requires_grad_(with_backward)
    ctx.requires_grad_(with_backward)
    for t in (den, sem, feat, rgb):
        t.requires_grad_(with_backward)
    with torch.set_grad_enabled(with_backward):
        vox = bb.get_voxel_feats(depth.unsqueeze(2) * ctx.unsqueeze(3), 0, mats)
        rend = bb.volume_rendering_from_multiple_views(geom, den, sem, feat, rgb)
    names = ["rgb", "seg", "depth", "bev_rgb", "bev_seg", "bev_height", "voxel_density", "voxel_output"]
    big = lambda a: a.detach().numpy().reshape(-1)[::stride_big].copy()
    out["vox_strided" if stride_big > 1 else "vox"] = big(vox) if stride_big > 1 else vox.detach().numpy()
    out["vox_absmax"] = float(vox.detach().abs().max())
    for n, r in zip(names, rend):
        r = r.detach()
        key = "r_" + n
        if stride_big > 1 and r.numel() > 400000:
            out[key + "_strided"] = big(r)
        else:
            out[key] = r.numpy()
        out[key + "_absmax"] = float(r.abs().max())
    if with_backward:
        cots = synth.make_cotangents([vox.shape] + [r.shape for r in rend])
        loss_lift = (vox * cots[0]).sum()
        g_depth, g_ctx = torch.autograd.grad(loss_lift, [depth, ctx])
        loss_r = sum((r * c).sum() for r, c in zip(rend, cots[1:]))
        g_den, g_sem, g_feat, g_rgb, g_beta = torch.autograd.grad(loss_r, [den, sem, feat, rgb, bb.density.beta])
        gs = max(stride_big, 7)
        for n, g in (("g_depth", g_depth), ("g_ctx", g_ctx), ("g_den", g_den), ("g_sem", g_sem),
                     ("g_feat", g_feat), ("g_rgb", g_rgb)):
            out[n + "_strided"] = g.numpy().reshape(-1)[::gs].copy()
            out[n + "_absmax"] = float(g.abs().max())
        out["meta_grad_stride"] = gs
        out["g_beta"] = float(g_beta)
    os.makedirs(GOLDEN, exist_ok=True)
    path = os.path.join(GOLDEN, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"{name}: {os.path.getsize(path) / 1e6:.2f} MB in {time.time() - t0:.1f}s  "
          f"valid={out['lift_valid_count']} mask={out['render_mask_count']}")


CASES = {
    "mini_val": (MINI, 2, "val", "random", 1, True, False),
    "mini_stress": (MINI, 2, "stress", "surface", 1, True, False),
    "r50_val_digest": (R50_256x704, 1, "val", "surface", 4099, True, False),
    "r50_stress_digest": (R50_256x704, 1, "stress", "random", 4099, True, False),   # gradients too (round 2)
}


def main():
    only = sys.argv[1:]            # python -m oracle.gen_golden [case ...]: regenerate just these
    for name, a in CASES.items():
        if only and name not in only:
            continue
        torch.manual_seed(0)
        run_case(name, *a)


if __name__ == "__main__":
    main()
