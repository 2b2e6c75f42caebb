"""Seeded synthetic nuScenes-shaped inputs for the path.

No dataset is reachable from the build or GPU boxes, so tests, ``smoke()`` and ``bench.py`` all
draw their inputs here.  The camera rig mirrors how the reference's loader builds ``mats_dict``
(/root/reference/src/datasets/nusc_det_seg_dataset.py:118-146 ida, 149-175 bda, 489-513 sampling,
652-659 sensor2ego / intrinsics): six cameras at nuScenes-like yaw angles, 1600x900 sensors,
fx=fy~1266, val-mode resize/crop, identity bda -- plus train-mode and stress variants.

Everything is produced on the CPU with a ``torch.Generator`` seeded ``1234 + sample index`` so
the same bits are regenerated on every box.
"""
from __future__ import annotations

import math
from typing import Dict

import torch

from .config import PathConfig

_YAWS_DEG = (55.0, 0.0, -55.0, 110.0, 180.0, -110.0)
_SENSOR_H, _SENSOR_W = 900, 1600


def _gen(seed: int) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    return g


def _uniform(g, lo, hi, *shape):
    return lo + (hi - lo) * torch.rand(*shape, generator=g, dtype=torch.float64)


def _ida(resize: float, crop, flip: bool, rotate_deg: float) -> torch.Tensor:
    """Same construction as the loader's ``img_transform`` (dataset:118-146), matrices only."""
    rot = torch.eye(2) * resize
    tran = -torch.tensor([float(crop[0]), float(crop[1])])
    if flip:
        A = torch.tensor([[-1.0, 0.0], [0.0, 1.0]])
        b = torch.tensor([float(crop[2] - crop[0]), 0.0])
        rot = A.matmul(rot)
        tran = A.matmul(tran) + b
    h = rotate_deg / 180.0 * math.pi
    A = torch.tensor([[math.cos(h), math.sin(h)], [-math.sin(h), math.cos(h)]])
    b = torch.tensor([float(crop[2] - crop[0]), float(crop[3] - crop[1])]) / 2
    b = A.matmul(-b) + b
    rot = A.matmul(rot)
    tran = A.matmul(tran) + b
    m = torch.zeros(4, 4)
    m[3, 3] = 1
    m[2, 2] = 1
    m[:2, :2] = rot
    m[:2, 3] = tran
    return m


def make_mats(cfg: PathConfig, batch: int, mode: str = "val", seed: int = 1234,
              with_bda: bool = True) -> Dict[str, torch.Tensor]:
    """``mats_dict`` as the reference's collate emits it (dataset:1014-1019).

    mode: 'val'    deterministic resize/crop, identity bda            (dataset:489-498)
          'train'  random resize in the config's resize_lim, random crop, rot 0  (base_exp.py:93-111)
          'stress' adds ida rotation +-5.4 deg and flips, bda rotation +-22.5 deg, scale, flips
    """
    fH, fW = cfg.final_dim
    N = cfg.num_cams
    s2e = torch.zeros(batch, 1, N, 4, 4)
    intr = torch.zeros(batch, 1, N, 4, 4)
    ida = torch.zeros(batch, 1, N, 4, 4)
    bda = torch.zeros(batch, 4, 4)
    cam_axes = torch.tensor([[0.0, 0.0, 1.0], [-1.0, 0.0, 0.0], [0.0, -1.0, 0.0]], dtype=torch.float64)
    base_resize = max(fH / _SENSOR_H, fW / _SENSOR_W)
    for b in range(batch):
        g = _gen(seed + b)
        for n in range(N):
            yaw = math.radians(_YAWS_DEG[n % len(_YAWS_DEG)])
            Rz = torch.tensor([[math.cos(yaw), -math.sin(yaw), 0.0],
                               [math.sin(yaw), math.cos(yaw), 0.0],
                               [0.0, 0.0, 1.0]], dtype=torch.float64)
            E = torch.eye(4, dtype=torch.float64)
            E[:3, :3] = Rz @ cam_axes
            E[:3, 3] = torch.tensor([1.5 * math.cos(yaw), 1.5 * math.sin(yaw), 1.5], dtype=torch.float64) \
                + _uniform(g, -0.1, 0.1, 3)
            s2e[b, 0, n] = E.float()
            Kmat = torch.zeros(4, 4, dtype=torch.float64)
            f = 1266.0 + _uniform(g, -10, 10, 1).item()
            Kmat[0, 0] = f
            Kmat[1, 1] = f
            Kmat[0, 2] = 816.0 + _uniform(g, -10, 10, 1).item()
            Kmat[1, 2] = 491.0 + _uniform(g, -10, 10, 1).item()
            Kmat[2, 2] = 1.0
            Kmat[3, 3] = 1.0
            intr[b, 0, n] = Kmat.float()
            if mode == "val":
                resize = base_resize
                newW, newH = int(_SENSOR_W * resize), int(_SENSOR_H * resize)
                crop_h = newH - fH
                crop_w = int(max(0, newW - fW) / 2)
                flip, rot = False, 0.0
            else:
                # train-mode resize_lim (0.386, 0.55) is quoted for 256x704; scale it with final_dim
                k = base_resize / 0.44
                resize = _uniform(g, 0.386 * k, 0.55 * k, 1).item()
                newW, newH = int(_SENSOR_W * resize), int(_SENSOR_H * resize)
                crop_h = newH - fH
                crop_w = int(_uniform(g, 0, max(0, newW - fW), 1).item())
                flip, rot = False, 0.0
                if mode == "stress":
                    flip = bool(torch.rand(1, generator=g).item() < 0.5)
                    rot = _uniform(g, -5.4, 5.4, 1).item()
            crop = (crop_w, crop_h, crop_w + fW, crop_h + fH)
            ida[b, 0, n] = _ida(resize, crop, flip, rot)
        if mode == "stress":
            ang = math.radians(_uniform(g, -22.5, 22.5, 1).item())
            sc = _uniform(g, 0.95, 1.05, 1).item()
            fdx = bool(torch.rand(1, generator=g).item() < 0.5)
            fdy = bool(torch.rand(1, generator=g).item() < 0.5)
            rot_mat = torch.tensor([[math.cos(ang), -math.sin(ang), 0.0],
                                    [math.sin(ang), math.cos(ang), 0.0], [0.0, 0.0, 1.0]])
            flip_mat = torch.eye(3)
            if fdx:
                flip_mat = flip_mat @ torch.tensor([[-1.0, 0, 0], [0, 1.0, 0], [0, 0, 1.0]])
            if fdy:
                flip_mat = flip_mat @ torch.tensor([[1.0, 0, 0], [0, -1.0, 0], [0, 0, 1.0]])
            m = torch.eye(4)
            m[:3, :3] = flip_mat @ (torch.eye(3) * sc @ rot_mat)
            bda[b] = m
        else:
            bda[b] = torch.eye(4)
    out = {"sensor2ego_mats": s2e, "intrin_mats": intr, "ida_mats": ida}
    if with_bda:
        out["bda_mat"] = bda
    return out


def make_lift_inputs(cfg: PathConfig, batch: int, seed: int = 1234, dtype=torch.float32):
    """depth = softmax(2*N(0,1)) over D (B,N,D,fH,fW); ctx = N(0,1) (B,N,C,fH,fW)."""
    N, D, C, fH, fW = cfg.num_cams, cfg.D, cfg.C, cfg.fH, cfg.fW
    depth = torch.empty(batch, N, D, fH, fW)
    ctx = torch.empty(batch, N, C, fH, fW)
    for b in range(batch):
        g = _gen(seed + 1000 + b)
        depth[b] = torch.softmax(2.0 * torch.randn(N, D, fH, fW, generator=g), dim=1)
        ctx[b] = torch.randn(N, C, fH, fW, generator=g)
    return depth.to(dtype), ctx.to(dtype)


def _surface_sdf(cfg: PathConfig, g: torch.Generator) -> torch.Tensor:
    """SDF-like field: ground plane at z=-1.5 m plus 20 random boxes; value = signed distance
    (negative inside), shifted by the density bias so that surfaces sit at s = sdf_bias."""
    from .lattice import build_lattice
    lat = build_lattice(cfg)
    zz, yy, xx = torch.meshgrid(lat.zs, lat.ys, lat.xs, indexing="ij")
    sdf = zz - (-1.5)
    for _ in range(20):
        c = torch.stack([_uniform(g, -45, 45, 1), _uniform(g, -45, 45, 1), _uniform(g, -1.5, 0.0, 1)]).float().view(3)
        h = torch.stack([_uniform(g, 0.8, 4.0, 1), _uniform(g, 0.8, 4.0, 1), _uniform(g, 0.6, 2.0, 1)]).float().view(3)
        q = torch.stack([(xx - c[0]).abs() - h[0], (yy - c[1]).abs() - h[1], (zz - c[2]).abs() - h[2]])
        outside = q.clamp(min=0).pow(2).sum(0).sqrt()
        inside = q.max(dim=0).values.clamp(max=0)
        sdf = torch.minimum(sdf, outside + inside)
    return sdf + cfg.sdf_bias


def make_render_inputs(cfg: PathConfig, batch: int, seed: int = 1234, field: str = "surface",
                       dtype=torch.float32):
    """The four volumes the render consumes (BV2:391): density_feature (B,1,Z,Y,X),
    semantic_logits (B,K,Z,Y,X), base_features (B,C,Z,Y,X), rgb (B,3,Z,Y,X).

    field: 'random'   density_feature = -1 + N(0.3, 0.5): rays saturate within ~10 m
           'surface'  SDF of a ground plane + boxes (+ 0.05 noise): realistic termination
           'empty'    density_feature = +1: sigma ~ 1e-9, rays never terminate (worst case)
    """
    Z, Y, X = cfg.vZ, cfg.vY, cfg.vX
    den = torch.empty(batch, 1, Z, Y, X)
    sem = torch.empty(batch, cfg.K, Z, Y, X)
    feat = torch.empty(batch, cfg.C, Z, Y, X)
    rgb = torch.empty(batch, 3, Z, Y, X)
    for b in range(batch):
        g = _gen(seed + 2000 + b)
        if field == "random":
            den[b, 0] = cfg.sdf_bias + 0.3 + 0.5 * torch.randn(Z, Y, X, generator=g)
        elif field == "surface":
            den[b, 0] = _surface_sdf(cfg, g) + 0.05 * torch.randn(Z, Y, X, generator=g)
        elif field == "empty":
            den[b, 0] = 1.0
        else:
            raise ValueError(field)
        sem[b] = torch.randn(cfg.K, Z, Y, X, generator=g)
        rgb[b] = torch.rand(3, Z, Y, X, generator=g)
        feat[b] = torch.randn(cfg.C, Z, Y, X, generator=g)
    return den.to(dtype), sem.to(dtype), feat.to(dtype), rgb.to(dtype)


def make_cotangents(shapes, seed: int = 1234):
    g = _gen(seed + 3000)
    return [torch.randn(*s, generator=g) for s in shapes]
