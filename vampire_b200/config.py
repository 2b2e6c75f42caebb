"""Configuration of the 2D->3D feature path.

Field names are the reference's ``backbone_conf`` keys
(/root/reference/src/exps/nuscenes/base_exp.py:40-92) so an experiment file can pass its
dict straight through, exactly as it does to ``BaseVAMPIRE2.__init__``
(/root/reference/src/layers/backbones/base_vampire2.py:83-104).
"""
from __future__ import annotations

from dataclasses import dataclass, asdict
from typing import Sequence, Tuple


def _count(bound: Sequence[float]) -> int:
    # the reference sizes every lattice axis with int((hi - lo) / step)  (BV2:276-281)
    return int((bound[1] - bound[0]) / bound[2])


@dataclass(frozen=True)
class PathConfig:
    x_bound_seg: Tuple[float, float, float] = (-51.2, 51.2, 0.4)
    y_bound_seg: Tuple[float, float, float] = (-51.2, 51.2, 0.4)
    z_bound_seg: Tuple[float, float, float] = (-5.0, 3.0, 0.4)
    x_bound_det: Tuple[float, float, float] = (-51.2, 51.2, 0.4)
    y_bound_det: Tuple[float, float, float] = (-51.2, 51.2, 0.4)
    z_bound_det: Tuple[float, float, float] = (-1.0, 3.0, 0.4)
    d_bound: Tuple[float, float, float] = (2.0, 70.4, 0.8)
    final_dim: Tuple[int, int] = (256, 704)
    downsample_factor: int = 4
    upsample_factor: int = 4
    mid_channels: int = 16
    num_classes: int = 18
    density_mode: str = "sdf"
    sdf_bias: float = -1.0
    cat_seg: bool = False
    num_cams: int = 6

    # ---- derived sizes -------------------------------------------------------------------
    @property
    def fH(self) -> int:
        return self.final_dim[0] // self.downsample_factor

    @property
    def fW(self) -> int:
        return self.final_dim[1] // self.downsample_factor

    @property
    def D(self) -> int:
        """Depth planes: len(torch.arange(*d_bound)) (BV2:258). Computed by the lattice."""
        import torch
        return int(torch.arange(*self.d_bound, dtype=torch.float).numel())

    @property
    def S(self) -> int:
        """Ray samples = D - 1 (BV2:148, 397)."""
        return self.D - 1

    @property
    def vZ(self) -> int:
        return _count(self.z_bound_seg)

    @property
    def vY(self) -> int:
        return _count(self.y_bound_seg)

    @property
    def vX(self) -> int:
        return _count(self.x_bound_seg)

    @property
    def oZ(self) -> int:
        return _count(self.z_bound_det)

    @property
    def oY(self) -> int:
        return _count(self.y_bound_det)

    @property
    def oX(self) -> int:
        return _count(self.x_bound_det)

    @property
    def C(self) -> int:
        return self.mid_channels

    @property
    def K(self) -> int:
        return self.num_classes

    @property
    def cam_channels(self) -> int:
        """density + semantics + rgb: the channels the camera branch consumes (BV2:423-425)."""
        return 1 + self.num_classes + 3

    @property
    def all_channels(self) -> int:
        """cat([density, sem, rgb, feat]) (BV2:396)."""
        return self.cam_channels + self.mid_channels

    def backbone_kwargs(self) -> dict:
        d = asdict(self)
        d.pop("num_cams")
        return d

    @staticmethod
    def from_backbone_conf(conf: dict, num_cams: int = 6) -> "PathConfig":
        keys = PathConfig.__dataclass_fields__.keys()
        kw = {k: (tuple(v) if isinstance(v, (list, tuple)) else v) for k, v in conf.items() if k in keys}
        kw["num_cams"] = num_cams
        return PathConfig(**kw)


# The target experiment (base_exp.py:40-92): 6 x 256x704, D=86, grid 20x256x256, C=16, K=18.
R50_256x704 = PathConfig()

# BASELINE.json configs[3]: scaled frustum 6 x 512x1408, same grid.
R50_512x1408 = PathConfig(final_dim=(512, 1408))

# Reduced geometry used by the committed golden fixtures and the CPU-side unit tests
# (same channel counts, every axis shrunk so the reference finishes in well under a second).
MINI = PathConfig(
    x_bound_seg=(-51.2, 51.2, 3.2), y_bound_seg=(-51.2, 51.2, 3.2), z_bound_seg=(-5.0, 3.0, 0.8),
    x_bound_det=(-51.2, 51.2, 3.2), y_bound_det=(-51.2, 51.2, 3.2), z_bound_det=(-1.0, 3.0, 0.8),
    d_bound=(2.0, 58.0, 2.0), final_dim=(64, 176),
)

NAMED = {"r50_256x704": R50_256x704, "r50_512x1408": R50_512x1408, "mini": MINI}
