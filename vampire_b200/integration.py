"""Drop-in attachment to a reference backbone instance (``BaseVAMPIRE2`` / ``BaseLSSImpaintor``).

``attach(backbone)`` rebinds the four path methods of an existing reference backbone
(/root/reference/src/layers/backbones/base_vampire2.py:314, 351, 391, 483 -- identical in
base_lss_impaintor.py:316-521) to the B200 kernels, keeping every signature, so the reference's own
``_forward_single_sweep`` (BV2:518-649) runs unmodified:

    backbone.get_geometry / get_pixel / get_voxel_feats / volume_rendering_from_multiple_views

and adds the two fused entry points ``backbone.lift_pool`` / ``backbone.render``.

``attach(backbone, fused=True)`` additionally rebinds ``_forward_single_sweep`` itself to
:func:`fused_forward_single_sweep`, which reproduces BV2:518-649 statement for statement with lines 551-563 ->
``depth_softmax`` + ``lift_pool`` and 554-559 + 576-630 -> ``query_points`` / ``occupancy`` / ``render`` /
``upsample2d`` / the BEV ``tanh`` epilogue, so that the 372 MB frustum tensor and the 70 MB geometry tensor are
never materialised and the benchmarked kernels are what the reference's own ``forward`` reaches -- no edit of the
reference source.  :func:`fused_backbone_class` does the same as a subclass (the way the reference switches
variants in its experiment files, ...depth_semantic.py:203-209).

The learnable ``density.beta`` stays the backbone's own ``nn.Parameter`` (state-dict key
``density.beta`` unchanged, so published checkpoints load as before): the attached module reads it
by reference.
"""
from __future__ import annotations

import types

import torch
import torch.nn as nn

from .view_transform import LiftRenderB200, UpsampleB200

_CONF_KEYS = ("x_bound_seg", "y_bound_seg", "z_bound_seg", "x_bound_det", "y_bound_det", "z_bound_det", "d_bound",
              "final_dim", "downsample_factor", "upsample_factor", "mid_channels", "num_classes", "density_mode",
              "sdf_bias", "cat_seg")


def backbone_conf_of(backbone: nn.Module) -> dict:
    """Recover the path's config from the attributes the reference constructor stores (BV2:127-144)."""
    return {k: getattr(backbone, k) for k in _CONF_KEYS if hasattr(backbone, k)}


def fused_forward_single_sweep(self, sweep_index, sweep_imgs, mats_dict, inrange_pts=None):
    """Drop-in for ``BaseVAMPIRE2._forward_single_sweep`` (BV2:518-649; twin base_lss_impaintor.py:528-655) on an
    attached backbone: same arguments, same 12-tuple, the 2D->3D path on the B200 kernels.  Everything that is not
    the path (image encoder, the two lift convs, the 3-D U-Net and heads, the BEV 1x1 conv) is the backbone's own
    modules, called exactly where the reference calls them."""
    path = self._vb200_path
    batch_size, num_sweeps, num_cams = sweep_imgs.shape[:3]
    img_feats = self.get_cam_feats(sweep_imgs)                                                        # BV2:549
    h, w = img_feats.shape[-2], img_feats.shape[-1]
    source_features = img_feats[:, 0, ...].reshape(batch_size * num_cams, -1, h, w)                  # BV2:550
    # BV2:551 .softmax(dim=1): fp32 arithmetic, fp32 output (autocast's behaviour for softmax)
    depth_softmax_features = path.depth_softmax(self.mapping_along_depth(source_features)).reshape(
        batch_size, num_cams, -1, h, w)
    low_channel_source_features = self.channel_lower(source_features).reshape(batch_size, num_cams, -1, h, w)  # :552
    # BV2:553 + 563: outer product + get_voxel_feats, never materialised
    voxel_features = path.lift_pool(depth_softmax_features, low_channel_source_features, mats_dict, sweep_index)
    if getattr(self, "cat_pos", False):                                                               # BV2:568-570
        norm_voxel_coords = self.norm_voxel_coords.permute(3, 0, 1, 2)[None, ...].repeat(batch_size, 1, 1, 1, 1)
        voxel_features = torch.cat([voxel_features, norm_voxel_coords.to(voxel_features.dtype)], dim=1)
    base_features = self.base_conv(voxel_features)                                                    # BV2:571
    density_feature = self.density_conv(base_features)                                                # BV2:573
    semantic_logits = self.seg_conv(base_features)                                                    # BV2:574
    rgb = self.rgb_conv(base_features)                                                                # BV2:575
    vols = [density_feature, semantic_logits, base_features, rgb]
    if len({t.dtype for t in vols}) > 1:     # e.g. a Sigmoid that autocast ran in fp32: one dtype for the kernels
        dt = torch.promote_types(torch.promote_types(vols[0].dtype, vols[1].dtype),
                                 torch.promote_types(vols[2].dtype, vols[3].dtype))
        density_feature, semantic_logits, base_features, rgb = (t.to(dt) for t in vols)

    pts_logits_batch, pts_sdf_batch = [], []                                                          # BV2:576-596
    if inrange_pts is not None:
        for i in range(batch_size):
            logits, sdf = path.query_points(semantic_logits[i:i + 1], density_feature[i:i + 1], inrange_pts[i])
            pts_logits_batch.append(logits[0])
            if self.density_mode == "sdf":
                pts_sdf_batch.append(sdf[0])
    # occupancy prediction                                                                           BV2:597-609
    if hasattr(self, "occ_coords"):
        occ_logits, occ_density = path.occupancy(semantic_logits, density_feature, mats_dict.get("bda_mat", None),
                                                 self.occ_coords)
    else:     # BaseLSSImpaintor: a fixed, unrotated grid (base_lss_impaintor.py:611-616)
        occ_logits, occ_density = path.occupancy(semantic_logits, density_feature, None, path.occ_coords())
    # BV2:554-559 + 612-614: geometry recomputed in-kernel (never stored), nan_to_num included.  Without gradients the
    # BEV epilogue of BV2:627-630 is folded into the BEV kernel (both operands are in registers there).
    fuse_tanh = not (torch.is_grad_enabled() and any(
        t.requires_grad for t in (density_feature, semantic_logits, base_features, rgb, path._beta())))
    (rgb_preds, seg_logits_preds, depth_preds, bev_rgb_preds, bev_seg_logits_preds, bev_height_preds, bev_density,
     voxel_output) = path.render(mats_dict, density_feature, semantic_logits, base_features, rgb, sweep_index,
                                 tanh_epilogue=fuse_tanh)
    up = self.upsample_factor
    fH, fW = self.fH, self.fW
    rgb_preds = path.upsample2d(rgb_preds.reshape(batch_size * num_cams, -1, fH, fW)).reshape(        # BV2:616-626
        batch_size, num_cams, -1, fH * up, fW * up)
    seg_logits_preds = path.upsample2d(seg_logits_preds.reshape(batch_size * num_cams, -1, fH, fW)).reshape(
        batch_size, num_cams, -1, fH * up, fW * up)
    depth_preds = path.upsample2d(depth_preds.reshape(batch_size * num_cams, -1, fH, fW)).reshape(
        batch_size, num_cams, -1, fH * up, fW * up)
    if not fuse_tanh:
        voxel_output = path.bev_epilogue(voxel_output, bev_density)                                   # BV2:627-630
    voxel_output_features = self.voxel_output(
        voxel_output.reshape(batch_size, -1, voxel_output.shape[-2], voxel_output.shape[-1])).float()  # BV2:631-632
    return (voxel_output_features.contiguous(), rgb_preds, seg_logits_preds, depth_preds, bev_rgb_preds,
            bev_seg_logits_preds, bev_height_preds, bev_density, pts_logits_batch, pts_sdf_batch, occ_logits,
            occ_density)


def fused_backbone_class(base_cls):
    """``class BaseVAMPIRE2B200(BaseVAMPIRE2)``: the reference backbone with the path on the B200
    kernels.  Same constructor, same state dict; selected in an experiment file by swapping the class."""

    class _Fused(base_cls):
        def __init__(self, *args, **kwargs):
            super().__init__(*args, **kwargs)
            attach(self, fused=True)

    _Fused.__name__ = base_cls.__name__ + "B200"
    _Fused.__qualname__ = _Fused.__name__
    return _Fused


def attach(backbone: nn.Module, channels_last_volume: bool = False, fused: bool = False,
           plans: str = "eval") -> nn.Module:
    conf = backbone_conf_of(backbone)
    path = LiftRenderB200(channels_last_volume=channels_last_volume, plans=plans, **conf)
    path.train(backbone.training)
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
    if fused:
        backbone._forward_single_sweep = types.MethodType(fused_forward_single_sweep, backbone)
    # `path` is not a registered submodule (the state dict must not grow): follow the backbone's train()/eval()
    # by hand, because the plan cache keys on it ("eval": plans only while not training)
    inner_train = backbone.train

    def train(self, mode: bool = True):
        self._vb200_path.train(mode)
        return inner_train(mode)

    backbone.train = types.MethodType(train, backbone)
    # the x4 upsample of the rendered maps right after the path (BV2:210, 616-626): parameter-free module
    if hasattr(backbone, "upsample2d") and hasattr(backbone, "upsample_factor"):
        backbone.upsample2d = UpsampleB200(backbone.upsample_factor)
    return backbone
