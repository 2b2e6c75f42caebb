"""Host-side mirror of the reference backbone's view-transform + render interface.

``LiftRenderB200`` exposes, with the reference's names, argument meaning and return shapes, the
methods of ``BaseVAMPIRE2`` that make up the 2D->3D feature path
(/root/reference/src/layers/backbones/base_vampire2.py):

    get_geometry(sensor2ego_mat, intrin_mat, ida_mat, bda_mat)                       BV2:314
    get_pixel(sensor2ego_mat, intrin_mat, ida_mat, bda_mat)                          BV2:351
    get_voxel_feats(frustum_feats, sweep_index, mats_dict)   [materialised frustum]  BV2:483
    volume_rendering_from_multiple_views(geom_xyz, density_feature, semantic_logits,
                                         voxel_features, rgb) -> 8-tuple            BV2:391

plus the two fused entry points the drop-in ``_forward_single_sweep`` uses instead of lines
BV2:553+563 and BV2:554-559+612-614:

    lift_pool(depth_softmax_features, low_channel_source_features, mats_dict)
    render(mats_dict, density_feature, semantic_logits, voxel_features, rgb)

It owns the one learnable parameter of the path, ``density.beta`` (render_utils.py:7), under the
reference's state-dict key so published checkpoints load unchanged.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
import torch.nn as nn

from . import cabi, ops
from .config import PathConfig
from .matrices import prepare_matrices
from .plan import PlanCache

Tensor = torch.Tensor


class LaplaceDensityParam(nn.Module):
    """Holds ``beta`` exactly like the reference's ``ModifyLaplaceDensity`` (render_utils.py:30-46);
    the density itself is evaluated inside the render kernels."""

    def __init__(self, beta: float = 0.1, bias: float = -1.0, beta_min: float = 1e-4):
        super().__init__()
        self.beta = nn.Parameter(torch.tensor(beta))
        self.beta_min = beta_min
        self.bias = bias

    def get_beta(self) -> Tensor:
        return self.beta.abs() + self.beta_min

    def forward(self, sdf: Tensor) -> Tensor:
        """render_utils.py:37-42, for callers outside the kernels (a reference ``_forward_single_sweep`` running on
        attached methods evaluates ``self.density(density_feature)`` itself for the occupancy query, BV2:609)."""
        beta = self.get_beta()
        x = sdf - self.bias
        return (1 / beta) * (0.5 + 0.5 * x.sign() * torch.expm1(-x.abs() / beta))


class UpsampleB200(nn.Module):
    """Drop-in for the reference's ``self.upsample2d = nn.UpsamplingBilinear2d(scale_factor=f)`` (BV2:210)."""

    def __init__(self, scale_factor: int):
        super().__init__()
        self.scale_factor = int(scale_factor)

    def forward(self, x: Tensor) -> Tensor:
        return ops.upsample_fwd(x, self.scale_factor)


class LiftRenderB200(nn.Module):
    def __init__(self, channels_last_volume: bool = False, plans: str = "eval", plan_cache_samples: int = 64,
                 **backbone_conf):
        """``backbone_conf``: the reference's dict (base_exp.py:40-92); unknown keys are ignored the
        way the image-encoder keys are irrelevant here.

        ``plans``: when to drive the lift from cached projection / sort plans (``vampire_b200.plan``):
        ``"eval"`` (default) while ``self.training`` is False -- validation / test matrices never change
        (nusc_det_seg_dataset.py:489-498, base_exp.py:113-120) -- ``"always"``, or ``"off"``.  Training draws a new
        ida every step (base_exp.py:93-111), so a plan would be rebuilt per call there."""
        super().__init__()
        if plans not in ("eval", "always", "off"):
            raise ValueError("plans must be 'eval', 'always' or 'off'")
        self.plans = plans
        self.plan_cache = PlanCache(plan_cache_samples)
        self._prep_cache = []      # [(key, (mats_dev, has_bda, mats_host))], most recent last
        self.cfg = PathConfig.from_backbone_conf(backbone_conf, backbone_conf.get("num_cams", 6))
        if self.cfg.C != 16 or self.cfg.K != 18 or not 1 <= self.cfg.num_cams <= 8:
            raise ValueError(f"libvb200 is compiled for mid_channels=16, num_classes=18 and at most 8 cameras "
                             f"(every reference experiment); got C={self.cfg.C}, K={self.cfg.K}, "
                             f"cams={self.cfg.num_cams}")
        if self.cfg.density_mode not in cabi.DENSITY_MODES:
            raise ValueError(f"density_mode must be 'sdf' or 'naive' (BV2:191-194), got {self.cfg.density_mode!r}")
        self.cfg_id = ops.register_config(self.cfg)
        # BV2:191-194: nn.Sigmoid() for 'naive' (the constructors' default), ModifyLaplaceDensity for 'sdf' (base_exp.py:51)
        self.density = (nn.Sigmoid() if self.cfg.density_mode == "naive"
                        else LaplaceDensityParam(beta=0.1, bias=self.cfg.sdf_bias))
        # the C ABI always takes a beta pointer; 'naive' never reads it
        self.register_buffer("_unit_beta", torch.ones(()), persistent=False)
        self.channels_last_volume = channels_last_volume
        self._det_points = None       # cat_seg: det-grid voxel centres, built on first use
        self._occ_pts = None          # occupancy(): the caller's Occ3D grid re-ordered x-fastest (identity + version keyed)
        st = ops.state(self.cfg_id)
        lat = st.lattice
        # the reference's buffers (BV2:146-160), rebuilt from the same 1-D axes
        self.register_buffer("camera_mids", lat.mids.clone(), persistent=False)
        self.register_buffer("bev_mids", lat.bev_mids.clone(), persistent=False)
        self.fD, self.fH, self.fW = self.cfg.S, self.cfg.fH, self.cfg.fW
        self.vZ, self.vY, self.vX = self.cfg.vZ, self.cfg.vY, self.cfg.vX

    # ---- matrices -----------------------------------------------------------------------------
    @staticmethod
    def _prep(sensor2ego_mat, intrin_mat, ida_mat, bda_mat, device) -> Tuple[Tensor, bool]:
        mats = prepare_matrices(sensor2ego_mat, intrin_mat, ida_mat, bda_mat)
        return mats.to(device, non_blocking=True), bda_mat is not None

    def _prep_dict(self, mats_dict: Dict[str, Tensor], sweep_index: int, device):
        mats, has_bda, _ = self._prep_dict_cached(mats_dict, sweep_index, device)
        return mats, has_bda

    def _prep_dict_cached(self, mats_dict: Dict[str, Tensor], sweep_index: int, device):
        """Prepared matrices of (mats_dict, sweep_index), computed once per distinct dict contents: the lift and
        the render of one ``_forward_single_sweep`` share them (three batched ``torch.inverse`` calls otherwise
        run twice per forward).  Returns (mats on `device`, has_bda, the same matrices on the host or None)."""
        device = torch.device(device)
        names = ("sensor2ego_mats", "intrin_mats", "ida_mats", "bda_mat")
        ts = [mats_dict.get(k, None) for k in names]
        # identity + version of the dict's tensors; the entry holds the tensors themselves, so an address can never
        # be recycled by a different tensor while it is cached
        vers = tuple(None if t is None else t._version for t in ts)
        for (k_sweep, k_dev, k_ts, k_vers), v in self._prep_cache:
            if k_sweep == sweep_index and k_dev == device and k_vers == vers and all(a is b for a, b in zip(k_ts, ts)):
                return v
        key = (sweep_index, device, ts, vers)
        prep = prepare_matrices(ts[0][:, sweep_index, ...], ts[1][:, sweep_index, ...], ts[2][:, sweep_index, ...], ts[3])
        host = prep if prep.device.type == "cpu" else None
        val = (prep.to(device, non_blocking=True), ts[3] is not None, host)
        self._prep_cache.append((key, val))
        del self._prep_cache[:-4]
        return val

    def _use_plans(self) -> bool:
        return self.plans == "always" or (self.plans == "eval" and not self.training)

    def _device(self) -> torch.device:
        beta = getattr(self.density, "beta", None)
        if beta is not None:
            return beta.device
        if self._unit_beta.device.type != "cuda" and torch.cuda.is_available():
            # 'naive' has no parameter to follow (an attached path is not a registered submodule): current device
            return torch.device("cuda", torch.cuda.current_device())
        return self._unit_beta.device

    def _beta(self, device: Optional[torch.device] = None) -> Tensor:
        """The learnable ``density.beta`` ('sdf'); a constant the kernels never read ('naive')."""
        beta = getattr(self.density, "beta", None)
        if beta is not None:
            return beta
        if device is not None and self._unit_beta.device != device:
            self._unit_beta = self._unit_beta.to(device)
        return self._unit_beta

    # ---- reference-named methods ----------------------------------------------------------------
    def get_geometry(self, sensor2ego_mat, intrin_mat, ida_mat, bda_mat) -> Tensor:
        mats, has_bda = self._prep(sensor2ego_mat, intrin_mat, ida_mat, bda_mat, self._device())
        return ops.get_geometry(mats, self.cfg_id, has_bda, False)

    def get_pixel(self, sensor2ego_mat, intrin_mat, ida_mat, bda_mat) -> Tensor:
        mats, has_bda = self._prep(sensor2ego_mat, intrin_mat, ida_mat, bda_mat, self._device())
        return ops.get_pixel(mats, self.cfg_id, has_bda)

    def get_voxel_feats(self, frustum_feats: Tensor, sweep_index: int, mats_dict: Dict[str, Tensor],
                        clamp_extreme: bool = True) -> Tensor:
        """Reference signature (BV2:483): the caller has already materialised the (B,N,C,D,fH,fW)
        frustum tensor.  Kept for drop-in compatibility; the fused ``lift_pool`` is the product path."""
        if not clamp_extreme:
            raise NotImplementedError("clamp_extreme=False is never used by the reference (BV2:563)")
        mats, has_bda = self._prep_dict(mats_dict, sweep_index, frustum_feats.device)
        out, _ = ops.gather_pool_fwd(frustum_feats, mats, self.cfg_id, has_bda)
        return out

    def volume_rendering_from_multiple_views(self, geom_xyz, density_feature, semantic_logits, voxel_features, rgb,
                                             mats_dict: Optional[Dict[str, Tensor]] = None):
        """Reference signature (BV2:391).  ``geom_xyz`` is consumed as given (the caller has applied
        nan_to_num, BV2:612).  Returns the reference's 8-tuple."""
        B = density_feature.shape[0]
        dev = density_feature.device
        # matrices are unused when geom is supplied; an identity block keeps the ABI uniform
        mats = torch.eye(4, device=dev).expand(B, self.cfg.num_cams, 6, 4, 4).contiguous()
        outs = ops.render_fwd(density_feature, semantic_logits, rgb, voxel_features, self._beta(dev), mats,
                              geom_xyz, self.cfg_id, True, cabi.BRANCH_CAM | cabi.BRANCH_BEV)
        return tuple(self._cat_seg(list(outs), semantic_logits))

    # ---- fused entry points -------------------------------------------------------------------
    def lift_pool(self, depth_softmax_features: Tensor, low_channel_source_features: Tensor,
                  mats_dict: Dict[str, Tensor], sweep_index: int = 0) -> Tensor:
        """BV2:553 + 563 without the (B,N,C,D,fH,fW) tensor.  depth (B,N,D,fH,fW), ctx (B,N,C,fH,fW)."""
        mats, has_bda, mats_host = self._prep_dict_cached(mats_dict, sweep_index, depth_softmax_features.device)
        depth, ctx = depth_softmax_features, low_channel_source_features
        # AMP (BV2:551-553): softmax is autocast to fp32, so an fp32 depth with 16-bit ctx is the reference's own
        # mix -- the product, the grid_sample and the pooled volume are fp32 there, and so they are here.  The only
        # other mix (16-bit depth, fp32 ctx) promotes the same way.
        if depth.dtype != ctx.dtype and depth.dtype != torch.float32:
            depth = depth.float()
        need_grad = torch.is_grad_enabled() and (depth.requires_grad or ctx.requires_grad)
        plan = None
        if self._use_plans():
            plan = self.plan_cache.lift(ops.state(self.cfg_id), self.cfg_id, mats, has_bda, mats_host).table
        out, _ = ops.lift_pool_fwd(depth, ctx, mats, self.cfg_id, has_bda, self.channels_last_volume, need_grad, plan)
        return out

    def lift_pool_2d(self, img_feats: Tensor, mats_dict: Dict[str, Tensor], sweep_index: int = 0) -> Tensor:
        """The ``BaseBiLinear`` ablation's lift (base_bilinear.py:471-517 ``get_voxel_feats``):
        no depth distribution -- every voxel centre with z > 0 bilinearly samples the (B,N,C,fH,fW) image
        features of each camera it projects into, non-zero mean over cameras.  It is the D = 1 case of the lift
        kernels (one depth plane of ones, depth test z > 0); gradients flow to ``img_feats``."""
        mats, has_bda, mats_host = self._prep_dict_cached(mats_dict, sweep_index, img_feats.device)
        if not hasattr(self, "_cfg_id_2d"):
            self._cfg_id_2d = ops.register_config(self.cfg, lift_2d=True)
        B, N, _, h, w = img_feats.shape
        ones = torch.ones(B, N, 1, h, w, dtype=img_feats.dtype, device=img_feats.device)
        need_grad = torch.is_grad_enabled() and img_feats.requires_grad
        plan = None
        if self._use_plans():
            plan = self.plan_cache.lift(ops.state(self._cfg_id_2d), self._cfg_id_2d, mats, has_bda, mats_host).table
        out, _ = ops.lift_pool_fwd(ones, img_feats, mats, self._cfg_id_2d, has_bda, self.channels_last_volume,
                                   need_grad, plan)
        return out

    def render(self, mats_dict: Dict[str, Tensor], density_feature: Tensor, semantic_logits: Tensor,
               voxel_features: Tensor, rgb: Tensor, sweep_index: int = 0, branches: int = 3,
               tanh_epilogue: bool = False):
        """BV2:554-559 + 612-614: geometry recomputed in-kernel from the matrices (never stored).
        ``tanh_epilogue=True`` (inference): the last output is ``voxel_output * bev_density.tanh()`` (BV2:627-630),
        formed inside the BEV kernel."""
        mats, has_bda, mats_host = self._prep_dict_cached(mats_dict, sweep_index, density_feature.device)
        plan = None
        if self._use_plans() and (branches & cabi.BRANCH_CAM):
            plan = self.plan_cache.render(ops.state(self.cfg_id), self.cfg_id, mats, has_bda, mats_host).table
        fuse = tanh_epilogue and not self.cfg.cat_seg        # the epilogue covers the concatenated seg channels too
        outs = list(ops.render_fwd(density_feature, semantic_logits, rgb, voxel_features, self._beta(density_feature.device), mats,
                                   None, self.cfg_id, has_bda, branches, plan, fuse))
        if branches & cabi.BRANCH_BEV:
            outs = self._cat_seg(outs, semantic_logits)
            if tanh_epilogue and not fuse:
                outs[7] = self.bev_epilogue(outs[7], outs[6])
        return tuple(outs)

    def det_points(self) -> Tensor:
        """(oZ*oY*oX, 3) ego coordinates of the det-grid voxel centres in the order of the BEV feature volume: the
        reference's ``output_coords`` (BV2:273-293, meshgrid ij order z, y, x) with z flipped (BV2:443), top level first."""
        lat = ops.state(self.cfg_id).lattice
        z, y, x = torch.meshgrid(lat.ozs.flip(0), lat.oys, lat.oxs, indexing="ij")
        return torch.stack([x, y, z], -1).reshape(-1, 3)

    def _cat_seg(self, outs, semantic_logits: Tensor):
        """``cat_seg=True`` (BV2:449-450; the default of BaseLSSImpaintor): the BEV feature volume carries the resampled,
        un-composited semantic logits behind the C feature channels.  They are the det-grid voxel centres sampled from
        the logits volume with zeros padding, top level first (BV2:442-443) -- the point-query kernel of SURVEY 8f row 3."""
        if not self.cfg.cat_seg:
            return outs
        cfg = self.cfg
        if self._det_points is None or self._det_points.device != semantic_logits.device:
            self._det_points = self.det_points().to(semantic_logits.device)
        seg, _ = ops.query_points_fwd(semantic_logits, self._det_points, None, None, self.cfg_id, False, False, False)
        seg = seg.reshape(seg.shape[0], cfg.K, cfg.oZ, cfg.oY, cfg.oX).to(outs[7].dtype)
        outs[7] = torch.cat([outs[7], seg], dim=1)
        return outs

    # ---- the callers right after the path ------------------------------------
    def depth_softmax(self, depth_logits: Tensor, out_fp32: bool = True) -> Tensor:
        """``mapping_along_depth(source_features).softmax(dim=1)`` (BV2:551) on (B*N, D, fH, fW) -- or
        (B, N, D, fH, fW) -- logits: softmax over the D depth planes in fp32.  ``out_fp32=True`` mirrors the
        reference under AMP (softmax is an autocast-to-fp32 op); ``False`` keeps the logits' dtype, which is what
        the bf16-feature lift consumes."""
        return ops.depth_softmax_fwd(depth_logits, out_fp32)

    def upsample2d(self, x: Tensor) -> Tensor:
        """``nn.UpsamplingBilinear2d(scale_factor=upsample_factor)`` (BV2:210) on (..., fH, fW) maps."""
        return ops.upsample_fwd(x, self.cfg.upsample_factor)

    def query_points(self, semantic_logits: Tensor, density_feature: Tensor, inrange_pts: Tensor):
        """LiDAR-point queries (BV2:578-596) for ONE sample's points (P,3) against every sample in the
        batch tensors given: returns (pts_logits (B,P,K) border-padded, pts_sdf (B,P) zero-padded * valid)."""
        logits, _ = ops.query_points_fwd(semantic_logits, inrange_pts, None, None, self.cfg_id, True, False, False)
        sdf, _ = ops.query_points_fwd(density_feature, inrange_pts, None, None, self.cfg_id, False, False, True)
        return logits.permute(0, 2, 1), sdf[:, 0]

    @staticmethod
    def occ_coords(point_cloud_range=(-40.0, -40.0, -1.0, 40.0, 40.0, 5.4), voxel=(0.4, 0.4, 0.4),
                   dims=(200, 200, 16)) -> Tensor:
        """Ego coordinates of the Occ3D grid, built with the reference's own torch CPU calls (BV2:295-302,
        ``create_norm_occ_coords(norm=False)``) -- for backbones that only keep the normalised copy."""
        idx = torch.where(torch.ones(dims, dtype=torch.bool))
        c = torch.cat([idx[a][:, None] * voxel[a] + voxel[a] / 2 + point_cloud_range[a] for a in range(3)], dim=1)
        return c.reshape(*dims, 3)

    def bev_epilogue(self, voxel_output: Tensor, bev_density: Tensor) -> Tensor:
        """``voxel_output * bev_density.tanh()`` for density_mode='sdf', ``* bev_density`` for 'naive' (BV2:627-630)."""
        scale = bev_density.tanh() if self.cfg.density_mode == "sdf" else bev_density
        return voxel_output * scale.to(voxel_output.dtype)

    def occupancy(self, semantic_logits: Tensor, density_feature: Tensor, bda_mat: Optional[Tensor],
                  occ_coords: Tensor):
        """Occ3D-grid queries (BV2:597-609, 647-648): occ_coords (X,Y,Z,3) ego coordinates of the
        200x200x16 grid, rotated per sample by bda[:3,:3] (``bda_mat=None``: the unrotated grid of
        ``BaseLSSImpaintor``, base_lss_impaintor.py:611-616).  Returns (occ_logits (B,X,Y,Z,K),
        tanh(occ_density) (B,X,Y,Z,1)) like the reference's return tuple."""
        # The grid arrives (X, Y, Z, 3) with z fastest; consecutive z are whole (y, x) planes apart in the volumes, so a
        # warp of consecutive points would gather from 32 planes.  Query it x-fastest instead (neighbouring threads,
        # neighbouring voxels) and hand the result back as the reference's (B, X, Y, Z, ch) view.
        X, Y, Z = occ_coords.shape[:-1]
        key = (occ_coords, occ_coords._version, semantic_logits.device)
        if self._occ_pts is None or self._occ_pts[0][0] is not key[0] or self._occ_pts[0][1:] != key[1:]:
            self._occ_pts = (key, occ_coords.to(semantic_logits.device).permute(2, 1, 0, 3).reshape(-1, 3).contiguous())
        pts = self._occ_pts[1]
        rot = None if bda_mat is None else bda_mat[:, :3, :3].to(semantic_logits.device)
        logits, _ = ops.query_points_fwd(semantic_logits, pts, rot, None, self.cfg_id, True, False, False)
        dens, _ = ops.query_points_fwd(density_feature, pts, rot, self._beta(density_feature.device), self.cfg_id, False, True, False)
        B = semantic_logits.shape[0]
        return (logits.reshape(B, -1, Z, Y, X).permute(0, 4, 3, 2, 1),
                dens.reshape(B, 1, Z, Y, X).permute(0, 4, 3, 2, 1).tanh())
