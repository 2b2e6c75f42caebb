"""Lattice tables of the path.

The reference builds its frustum / voxel-centre / mid-depth buffers once with torch CPU calls
(``create_frustum`` BV2:253-271, ``create_voxel_coords`` BV2:273-293, ``create_camera_mids``
BV2:243-246, ``create_bev_mids`` BV2:248-251).  Index parity needs the *same bits*: recomputing
``linspace``/``arange`` entries inside a kernel from the integer index is not bit-identical
(SURVEY B.8), so the kernels take these tables as 1-D fp32 arrays.  The meshgrid/stack the
reference performs only broadcasts the 1-D axes, hence the axes are all the kernels need.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch

from .config import PathConfig


def _centres(bound) -> torch.Tensor:
    lo, hi, step = bound
    return torch.linspace(lo + step / 2.0, hi - step / 2.0, int((hi - lo) / step), dtype=torch.float)


@dataclass
class Lattice:
    us: torch.Tensor        # (fW,)  image-x of feature columns   BV2:262
    vs: torch.Tensor        # (fH,)  image-y of feature rows      BV2:264
    ds: torch.Tensor        # (D,)   depth planes                 BV2:258
    xs: torch.Tensor        # (vX,)  seg-grid voxel centres       BV2:280
    ys: torch.Tensor        # (vY,)
    zs: torch.Tensor        # (vZ,)
    oxs: torch.Tensor       # (oX,)  det/BEV-grid voxel centres   BV2:160
    oys: torch.Tensor       # (oY,)
    ozs: torch.Tensor       # (oZ,)
    mids: torch.Tensor      # (S,)   interval mid depths          BV2:243-246
    bev_mids: torch.Tensor  # (oZ,)  BEV level heights, top first BV2:248-251

    def to(self, device) -> "Lattice":
        return Lattice(**{k: v.to(device) for k, v in self.__dict__.items()})

    def packed(self) -> torch.Tensor:
        """All tables in one contiguous fp32 buffer, order = field order (see csrc VbTables)."""
        return torch.cat([v.reshape(-1) for v in self.__dict__.values()]).contiguous()


def build_lattice(cfg: PathConfig) -> Lattice:
    ogfH, ogfW = cfg.final_dim
    ds = torch.arange(*cfg.d_bound, dtype=torch.float)
    return Lattice(
        us=torch.linspace(0, ogfW - 1, cfg.fW, dtype=torch.float),
        vs=torch.linspace(0, ogfH - 1, cfg.fH, dtype=torch.float),
        ds=ds,
        xs=_centres(cfg.x_bound_seg), ys=_centres(cfg.y_bound_seg), zs=_centres(cfg.z_bound_seg),
        oxs=_centres(cfg.x_bound_det), oys=_centres(cfg.y_bound_det), ozs=_centres(cfg.z_bound_det),
        mids=0.5 * (ds[:-1] + ds[1:]),
        bev_mids=torch.flip(_centres(cfg.z_bound_det), dims=[0]),
    )
