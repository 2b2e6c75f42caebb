"""Drop-in attachment to a reference backbone instance (``BaseVAMPIRE2`` / ``BaseLSSImpaintor``).

``attach(backbone)`` rebinds the four path methods of an existing reference backbone
(/root/reference/src/layers/backbones/base_vampire2.py:314, 351, 391, 483 -- identical in
base_lss_impaintor.py:316-521) to the B200 kernels, keeping every signature, so the reference's own
``_forward_single_sweep`` (BV2:518-649) runs unmodified:

    backbone.get_geometry / get_pixel / get_voxel_feats / volume_rendering_from_multiple_views

and adds the two fused entry points ``backbone.lift_pool`` / ``backbone.render`` that a two-line
edit of ``_forward_single_sweep`` switches to (INTEGRATION.md shows the diff) so that the 372 MB
frustum tensor and the 70 MB geometry tensor are never materialised.

The learnable ``density.beta`` stays the backbone's own ``nn.Parameter`` (state-dict key
``density.beta`` unchanged, so published checkpoints load as before): the attached module reads it
by reference.
"""
from __future__ import annotations

import types

import torch.nn as nn

from .view_transform import LiftRenderB200, UpsampleB200

_CONF_KEYS = ("x_bound_seg", "y_bound_seg", "z_bound_seg", "x_bound_det", "y_bound_det", "z_bound_det", "d_bound",
              "final_dim", "downsample_factor", "upsample_factor", "mid_channels", "num_classes", "density_mode",
              "sdf_bias", "cat_seg")


def backbone_conf_of(backbone: nn.Module) -> dict:
    """Recover the path's config from the attributes the reference constructor stores (BV2:127-144)."""
    return {k: getattr(backbone, k) for k in _CONF_KEYS if hasattr(backbone, k)}


def attach(backbone: nn.Module, channels_last_volume: bool = False) -> nn.Module:
    conf = backbone_conf_of(backbone)
    path = LiftRenderB200(channels_last_volume=channels_last_volume, **conf)
    # share the reference's parameter object instead of owning a copy
    path.density = backbone.density
    # keep `path` out of backbone._modules so the state dict does not grow new keys
    object.__setattr__(backbone, "_vb200_path", path)

    def get_geometry(self, sensor2ego_mat, intrin_mat, ida_mat, bda_mat):
        return self._vb200_path.get_geometry(sensor2ego_mat, intrin_mat, ida_mat, bda_mat)

    def get_pixel(self, sensor2ego_mat, intrin_mat, ida_mat, bda_mat):
        return self._vb200_path.get_pixel(sensor2ego_mat, intrin_mat, ida_mat, bda_mat)

    def get_voxel_feats(self, frustum_feats, sweep_index, mats_dict, clamp_extreme=True):
        if frustum_feats.dim() == 5:
            # BaseBiLinear.get_voxel_feats (base_bilinear.py:471): (B,N,C,h,w) image features, 2-D lift
            if not clamp_extreme:
                raise NotImplementedError("clamp_extreme=False is never used by the reference")
            return self._vb200_path.lift_pool_2d(frustum_feats, mats_dict, sweep_index)
        return self._vb200_path.get_voxel_feats(frustum_feats, sweep_index, mats_dict, clamp_extreme)

    def volume_rendering_from_multiple_views(self, geom_xyz, density_feature, semantic_logits, voxel_features, rgb):
        return self._vb200_path.volume_rendering_from_multiple_views(geom_xyz, density_feature, semantic_logits,
                                                                     voxel_features, rgb)

    def lift_pool(self, depth_softmax_features, low_channel_source_features, mats_dict, sweep_index=0):
        return self._vb200_path.lift_pool(depth_softmax_features, low_channel_source_features, mats_dict, sweep_index)

    def render(self, mats_dict, density_feature, semantic_logits, voxel_features, rgb, sweep_index=0):
        return self._vb200_path.render(mats_dict, density_feature, semantic_logits, voxel_features, rgb, sweep_index)

    def depth_softmax(self, depth_logits, out_fp32=True):
        return self._vb200_path.depth_softmax(depth_logits, out_fp32)

    for fn in (get_geometry, get_pixel, get_voxel_feats, volume_rendering_from_multiple_views, lift_pool, render,
               depth_softmax):
        setattr(backbone, fn.__name__, types.MethodType(fn, backbone))
    # the x4 upsample of the rendered maps right after the path (BV2:210, 616-626): parameter-free module
    if hasattr(backbone, "upsample2d") and hasattr(backbone, "upsample_factor"):
        backbone.upsample2d = UpsampleB200(backbone.upsample_factor)
    return backbone
