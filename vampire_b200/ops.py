"""torch.library custom ops over the extern "C" entry points of libvb200.so.

Each op body does nothing but allocate outputs/workspace from the PyTorch caching allocator,
fetch ``data_ptr()`` / the current CUDA stream and call the C ABI through ctypes
(``vampire_b200.cabi``).  Autograd is wired with ``register_autograd`` to the matching backward
entry points.  There is no CPU implementation: a CPU tensor raises.

Static configuration (sizes, fp32 constants, lattice tables) cannot travel through an op schema,
so a :class:`PathState` is registered once per config and ops receive its integer handle.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Tuple

import torch

from . import cabi
from .config import PathConfig
from .lattice import Lattice, build_lattice

Tensor = torch.Tensor


class PathState:
    """Per-config state: the lattice (host) and its per-device copies."""

    def __init__(self, cfg: PathConfig, lift_2d: bool = False):
        self.cfg = cfg
        # the BaseBiLinear ablation's 2-D lift (base_bilinear.py:471-517): same kernels, one depth plane of ones
        self.lift_2d = lift_2d
        self.lattice: Lattice = build_lattice(cfg)
        self._tables: Dict[torch.device, cabi.DeviceTables] = {}
        self.term_eps = 1e-8
        # samples packed+marched per round of the render: 0 = the whole batch in one round (best
        # occupancy, measured 1.8x faster at B=8); 1 = sample by sample (packed copy stays L2-resident)
        self.render_group = 0

    def tables(self, device: torch.device) -> cabi.DeviceTables:
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("vampire_b200 ops run on CUDA (sm_100a) tensors only; there is no CPU path")
        if device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        if device not in self._tables:
            self._tables[device] = cabi.DeviceTables(self.lattice, device)
        return self._tables[device]

    @property
    def lift_D(self) -> int:
        """Depth planes of the lift input: the config's D, or 1 for the 2-D lift."""
        return 1 if self.lift_2d else self.cfg.D

    def grid(self, batch: int, has_bda: bool) -> cabi.VbGrid:
        return cabi.make_grid(self.cfg, batch, has_bda, self.term_eps, self.lift_2d)


_STATES: List[PathState] = []
_BY_CFG: Dict[Tuple[PathConfig, bool], int] = {}


def register_config(cfg: PathConfig, lift_2d: bool = False) -> int:
    """Handle of a path configuration.  ``lift_2d=True`` registers the variant whose lift ops implement the
    BaseBiLinear 2-D lift (only ``lift_pool_fwd/bwd`` and ``lift_indices`` are meaningful on it)."""
    key = (cfg, bool(lift_2d))
    if key not in _BY_CFG:
        _STATES.append(PathState(cfg, bool(lift_2d)))
        _BY_CFG[key] = len(_STATES) - 1
    return _BY_CFG[key]


def state(handle: int) -> PathState:
    return _STATES[handle]


def _need_cuda(*ts: Optional[Tensor]) -> torch.device:
    dev = None
    for t in ts:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("vampire_b200: expected CUDA tensors (no CPU fallback exists for this path)")
        dev = t.device if dev is None else dev
        if t.device != dev:
            raise RuntimeError("vampire_b200: tensors on different devices")
    return dev


def _mats_ok(mats: Tensor, B: int, N: int) -> Tensor:
    if mats.shape != (B, N, 6, 4, 4) or mats.dtype != torch.float32:
        raise ValueError(f"mats must be (B={B}, N={N}, 6, 4, 4) fp32 from prepare_matrices, got {tuple(mats.shape)}")
    return mats.contiguous()


# =============================================================================================
# geometry (no autograd: the reference computes it from constants and matrices only)
# =============================================================================================
@torch.library.custom_op("vampire_b200::get_pixel", mutates_args=())
def get_pixel(mats: Tensor, cfg_id: int, has_bda: bool) -> Tensor:
    st = state(cfg_id)
    cfg = st.cfg
    dev = _need_cuda(mats)
    B = mats.shape[0]
    mats = _mats_ok(mats, B, cfg.num_cams)
    out = torch.empty(B, cfg.num_cams, cfg.vZ, cfg.vY, cfg.vX, 3, dtype=torch.float32, device=dev)
    g = st.grid(B, has_bda)
    with torch.cuda.device(dev):
        cabi.check(cabi.lib().vb200_get_pixel(C.byref(g), C.byref(st.tables(dev).struct), mats.data_ptr(),
                                              out.data_ptr(), cabi.stream_ptr(dev)))
    return out


@get_pixel.register_fake
def _(mats, cfg_id, has_bda):
    cfg = state(cfg_id).cfg
    return mats.new_empty(mats.shape[0], cfg.num_cams, cfg.vZ, cfg.vY, cfg.vX, 3)


@torch.library.custom_op("vampire_b200::get_geometry", mutates_args=())
def get_geometry(mats: Tensor, cfg_id: int, has_bda: bool, nan_to_num: bool) -> Tensor:
    st = state(cfg_id)
    cfg = st.cfg
    dev = _need_cuda(mats)
    B = mats.shape[0]
    mats = _mats_ok(mats, B, cfg.num_cams)
    out = torch.empty(B, cfg.num_cams, cfg.D, cfg.fH, cfg.fW, 3, dtype=torch.float32, device=dev)
    g = st.grid(B, has_bda)
    with torch.cuda.device(dev):
        cabi.check(cabi.lib().vb200_get_geometry(C.byref(g), C.byref(st.tables(dev).struct), mats.data_ptr(),
                                                 out.data_ptr(), int(nan_to_num), cabi.stream_ptr(dev)))
    return out


@get_geometry.register_fake
def _(mats, cfg_id, has_bda, nan_to_num):
    cfg = state(cfg_id).cfg
    return mats.new_empty(mats.shape[0], cfg.num_cams, cfg.D, cfg.fH, cfg.fW, 3)


def lift_indices(mats: Tensor, cfg_id: int, has_bda: bool) -> Tuple[Tensor, Tensor, Tensor]:
    """(valid uint8 (B,N,Z,Y,X), i0 int16 (...,3) = (x0,y0,z0), frac fp32 (...,3)) -- parity probe."""
    st = state(cfg_id)
    cfg = st.cfg
    dev = _need_cuda(mats)
    B = mats.shape[0]
    mats = _mats_ok(mats, B, cfg.num_cams)
    shp = (B, cfg.num_cams, cfg.vZ, cfg.vY, cfg.vX)
    valid = torch.empty(shp, dtype=torch.uint8, device=dev)
    i0 = torch.empty(shp + (3,), dtype=torch.int16, device=dev)
    frac = torch.empty(shp + (3,), dtype=torch.float32, device=dev)
    g = st.grid(B, has_bda)
    with torch.cuda.device(dev):
        cabi.check(cabi.lib().vb200_lift_indices(C.byref(g), C.byref(st.tables(dev).struct), mats.data_ptr(),
                                                 valid.data_ptr(), i0.data_ptr(), frac.data_ptr(),
                                                 cabi.stream_ptr(dev)))
    return valid, i0, frac


def render_indices(mats: Tensor, cfg_id: int, has_bda: bool, geom: Optional[Tensor] = None):
    """(mask uint8 (B,N,S,fH,fW), i0 int16 (...,3), frac fp32 (...,3)) -- parity probe."""
    st = state(cfg_id)
    cfg = st.cfg
    dev = _need_cuda(mats, geom)
    B = mats.shape[0]
    mats = _mats_ok(mats, B, cfg.num_cams)
    shp = (B, cfg.num_cams, cfg.S, cfg.fH, cfg.fW)
    mask = torch.empty(shp, dtype=torch.uint8, device=dev)
    i0 = torch.empty(shp + (3,), dtype=torch.int16, device=dev)
    frac = torch.empty(shp + (3,), dtype=torch.float32, device=dev)
    if geom is not None:
        geom = geom.float().contiguous()
    g = st.grid(B, has_bda)
    with torch.cuda.device(dev):
        cabi.check(cabi.lib().vb200_render_indices(C.byref(g), C.byref(st.tables(dev).struct), mats.data_ptr(),
                                                   cabi.ptr(geom), mask.data_ptr(), i0.data_ptr(), frac.data_ptr(),
                                                   cabi.stream_ptr(dev)))
    return mask, i0, frac


# =============================================================================================
# lift + pool
# =============================================================================================
def _lift_dtypes(depth: Tensor, ctx: Tensor) -> Tuple[int, int]:
    """(dtype, ctx_dtype) codes: equal dtypes, or fp32 depth with 16-bit ctx -- the reference under AMP, where the
    fp32 softmax output promotes the frustum, the grid_sample and the pooled volume to fp32 (BV2:551-553)."""
    if depth.dtype != ctx.dtype and not (depth.dtype == torch.float32 and ctx.dtype in (torch.bfloat16, torch.float16)):
        raise TypeError(f"lift_pool: depth {depth.dtype} / ctx {ctx.dtype}: dtypes must match, or depth fp32 with "
                        f"16-bit ctx (AMP)")
    return cabi.dtype_code(depth.dtype), cabi.dtype_code(ctx.dtype)


def _plan_ok(plan: Optional[Tensor], B: int, dev) -> Optional[Tensor]:
    if plan is None:
        return None
    if plan.dtype != torch.int64 or plan.shape != (B, 4) or plan.device != dev or not plan.is_contiguous():
        raise ValueError("lift_pool: plan must be the (B, 4) int64 device table of a LiftPlanBatch")
    return plan


def lift_pool_fwd(depth: Tensor, ctx: Tensor, mats: Tensor, cfg_id: int, has_bda: bool, channels_last: bool,
                  save_cnt: bool, plan: Optional[Tensor] = None) -> Tuple[Tensor, Tensor]:
    """``plan``: the device table of a :class:`vampire_b200.plan.LiftPlanBatch` built for these matrices; the
    projection is then read from the cached plan instead of being recomputed (bit-identical result)."""
    return _lift_pool_fwd(depth, ctx, mats, cfg_id, has_bda, channels_last, save_cnt, plan)


# (the op schemas carry NO default values: the dispatcher strips arguments equal to their default before autograd
#  sees them, which would make the arity of the backward depend on whether a plan was passed)
@torch.library.custom_op("vampire_b200::lift_pool_fwd", mutates_args=())
def _lift_pool_fwd(depth: Tensor, ctx: Tensor, mats: Tensor, cfg_id: int, has_bda: bool, channels_last: bool,
                   save_cnt: bool, plan: Optional[Tensor]) -> Tuple[Tensor, Tensor]:
    st = state(cfg_id)
    cfg = st.cfg
    dev = _need_cuda(depth, ctx, mats, plan)
    B, N = depth.shape[:2]
    if depth.shape != (B, N, st.lift_D, cfg.fH, cfg.fW) or ctx.shape != (B, N, cfg.C, cfg.fH, cfg.fW):
        raise ValueError(f"lift_pool: depth {tuple(depth.shape)} / ctx {tuple(ctx.shape)} do not match the config")
    dt, cdt = _lift_dtypes(depth, ctx)
    mats = _mats_ok(mats, B, N)
    plan = _plan_ok(plan, B, dev)
    depth = depth.contiguous()
    ctx = ctx.contiguous()
    nvox = cfg.vZ * cfg.vY * cfg.vX
    if channels_last:
        out = torch.empty(B, cfg.vZ, cfg.vY, cfg.vX, cfg.C, dtype=depth.dtype, device=dev).permute(0, 4, 1, 2, 3)
    else:
        out = torch.empty(B, cfg.C, cfg.vZ, cfg.vY, cfg.vX, dtype=depth.dtype, device=dev)
    cnt = torch.empty(B, nvox if save_cnt else 0, dtype=torch.int64, device=dev)
    g = st.grid(B, has_bda)
    lib = cabi.lib()
    ws_bytes = lib.vb200_lift_pool_fwd_workspace(C.byref(g), cdt)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    layout = cabi.NDHWC if channels_last else cabi.NCDHW
    with torch.cuda.device(dev):
        if plan is None:
            cabi.check(lib.vb200_lift_pool_fwd(C.byref(g), C.byref(st.tables(dev).struct), mats.data_ptr(),
                                               depth.data_ptr(), ctx.data_ptr(), dt, cdt, out.data_ptr(), layout,
                                               cnt.data_ptr() if save_cnt else None, ws.data_ptr(), ws_bytes,
                                               cabi.stream_ptr(dev)))
        else:
            cabi.check(lib.vb200_lift_pool_fwd_planned(C.byref(g), plan.data_ptr(), depth.data_ptr(), ctx.data_ptr(),
                                                       dt, cdt, out.data_ptr(), layout,
                                                       cnt.data_ptr() if save_cnt else None, ws.data_ptr(), ws_bytes,
                                                       cabi.stream_ptr(dev)))
    return out, cnt


@_lift_pool_fwd.register_fake
def _(depth, ctx, mats, cfg_id, has_bda, channels_last, save_cnt, plan):
    cfg = state(cfg_id).cfg
    B = depth.shape[0]
    if channels_last:
        out = depth.new_empty(B, cfg.vZ, cfg.vY, cfg.vX, cfg.C).permute(0, 4, 1, 2, 3)
    else:
        out = depth.new_empty(B, cfg.C, cfg.vZ, cfg.vY, cfg.vX)
    cnt = depth.new_empty(B, cfg.vZ * cfg.vY * cfg.vX if save_cnt else 0, dtype=torch.int64)
    return out, cnt


def lift_pool_bwd(gout: Tensor, depth: Tensor, ctx: Tensor, mats: Tensor, cnt: Tensor, cfg_id: int,
                  has_bda: bool, plan: Optional[Tensor] = None) -> Tuple[Tensor, Tensor]:
    return _lift_pool_bwd(gout, depth, ctx, mats, cnt, cfg_id, has_bda, plan)


@torch.library.custom_op("vampire_b200::lift_pool_bwd", mutates_args=())
def _lift_pool_bwd(gout: Tensor, depth: Tensor, ctx: Tensor, mats: Tensor, cnt: Tensor, cfg_id: int,
                   has_bda: bool, plan: Optional[Tensor]) -> Tuple[Tensor, Tensor]:
    st = state(cfg_id)
    cfg = st.cfg
    dev = _need_cuda(gout, depth, ctx, mats, cnt, plan)
    B, N = depth.shape[:2]
    dt, cdt = _lift_dtypes(depth, ctx)
    mats = _mats_ok(mats, B, N)
    plan = _plan_ok(plan, B, dev)
    depth = depth.contiguous()
    ctx = ctx.contiguous()
    gout = gout.to(depth.dtype)
    if gout.permute(0, 2, 3, 4, 1).is_contiguous():
        layout = cabi.NDHWC
    else:
        gout = gout.contiguous()
        layout = cabi.NCDHW
    gdepth = torch.empty_like(depth)
    gctx = torch.empty_like(ctx)
    g = st.grid(B, has_bda)
    lib = cabi.lib()
    if plan is None:
        ws_bytes = lib.vb200_lift_pool_bwd_workspace(C.byref(g), dt)
    else:
        ws_bytes = lib.vb200_lift_pool_bwd_planned_workspace(C.byref(g), dt)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        if plan is None:
            cabi.check(lib.vb200_lift_pool_bwd(C.byref(g), C.byref(st.tables(dev).struct), mats.data_ptr(),
                                               depth.data_ptr(), ctx.data_ptr(), dt, cdt, gout.data_ptr(), layout,
                                               cnt.data_ptr(), gdepth.data_ptr(), gctx.data_ptr(), ws.data_ptr(),
                                               ws_bytes, cabi.stream_ptr(dev)))
        else:
            cabi.check(lib.vb200_lift_pool_bwd_planned(C.byref(g), plan.data_ptr(), depth.data_ptr(), ctx.data_ptr(),
                                                       dt, cdt, gout.data_ptr(), layout, cnt.data_ptr(),
                                                       gdepth.data_ptr(), gctx.data_ptr(), ws.data_ptr(), ws_bytes,
                                                       cabi.stream_ptr(dev)))
    return gdepth, gctx


@_lift_pool_bwd.register_fake
def _(gout, depth, ctx, mats, cnt, cfg_id, has_bda, plan):
    return torch.empty_like(depth), torch.empty_like(ctx)


def _lift_setup(ctx, inputs, output):
    depth, context, mats, cfg_id, has_bda, channels_last, save_cnt, plan = inputs
    _, cnt = output
    ctx.save_for_backward(depth, context, mats, cnt, plan)
    ctx.cfg_id, ctx.has_bda, ctx.save_cnt = cfg_id, has_bda, save_cnt
    # the plan table holds raw pointers: keep the buffers they point into alive until the backward has run
    ctx.plan_keepalive = getattr(plan, "_vb200_keepalive", None) if plan is not None else None


def _lift_backward(ctx, gout, gcnt):
    if not ctx.save_cnt:
        raise RuntimeError("lift_pool_fwd was called with save_cnt=False: no backward possible")
    depth, context, mats, cnt, plan = ctx.saved_tensors
    gdepth, gctx = lift_pool_bwd(gout, depth, context, mats, cnt, ctx.cfg_id, ctx.has_bda, plan)
    return gdepth, gctx, None, None, None, None, None, None


torch.library.register_autograd("vampire_b200::lift_pool_fwd", _lift_backward, setup_context=_lift_setup)


# ---- compatibility: the reference's get_voxel_feats signature (materialised frustum tensor) --------
@torch.library.custom_op("vampire_b200::gather_pool_fwd", mutates_args=())
def gather_pool_fwd(frustum: Tensor, mats: Tensor, cfg_id: int, has_bda: bool) -> Tuple[Tensor, Tensor]:
    st = state(cfg_id)
    cfg = st.cfg
    dev = _need_cuda(frustum, mats)
    B, N = frustum.shape[:2]
    if frustum.shape != (B, N, cfg.C, cfg.D, cfg.fH, cfg.fW):
        raise ValueError(f"get_voxel_feats: frustum_feats has shape {tuple(frustum.shape)}")
    dt = cabi.dtype_code(frustum.dtype)
    mats = _mats_ok(mats, B, N)
    frustum = frustum.contiguous()
    out = torch.empty(B, cfg.C, cfg.vZ, cfg.vY, cfg.vX, dtype=frustum.dtype, device=dev)
    cnt = torch.empty(B, cfg.vZ * cfg.vY * cfg.vX, dtype=torch.int64, device=dev)
    g = st.grid(B, has_bda)
    with torch.cuda.device(dev):
        cabi.check(cabi.lib().vb200_gather_pool_fwd(C.byref(g), C.byref(st.tables(dev).struct), mats.data_ptr(),
                                                    frustum.data_ptr(), dt, out.data_ptr(), cnt.data_ptr(),
                                                    cabi.stream_ptr(dev)))
    return out, cnt


@gather_pool_fwd.register_fake
def _(frustum, mats, cfg_id, has_bda):
    cfg = state(cfg_id).cfg
    B = frustum.shape[0]
    return (frustum.new_empty(B, cfg.C, cfg.vZ, cfg.vY, cfg.vX),
            frustum.new_empty(B, cfg.vZ * cfg.vY * cfg.vX, dtype=torch.int64))


@torch.library.custom_op("vampire_b200::gather_pool_bwd", mutates_args=())
def gather_pool_bwd(gout: Tensor, mats: Tensor, cnt: Tensor, cfg_id: int, has_bda: bool) -> Tensor:
    st = state(cfg_id)
    cfg = st.cfg
    dev = _need_cuda(gout, mats, cnt)
    B, N = gout.shape[0], cfg.num_cams
    dt = cabi.dtype_code(gout.dtype)
    mats = _mats_ok(mats, B, N)
    gout = gout.contiguous()
    gfr = torch.empty(B, N, cfg.C, cfg.D, cfg.fH, cfg.fW, dtype=gout.dtype, device=dev)
    g = st.grid(B, has_bda)
    lib = cabi.lib()
    ws_bytes = lib.vb200_gather_pool_bwd_workspace(C.byref(g), dt)
    ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        cabi.check(lib.vb200_gather_pool_bwd(C.byref(g), C.byref(st.tables(dev).struct), mats.data_ptr(),
                                             gout.data_ptr(), cnt.data_ptr(), dt, gfr.data_ptr(), ws.data_ptr(),
                                             ws_bytes, cabi.stream_ptr(dev)))
    return gfr


@gather_pool_bwd.register_fake
def _(gout, mats, cnt, cfg_id, has_bda):
    cfg = state(cfg_id).cfg
    return gout.new_empty(gout.shape[0], cfg.num_cams, cfg.C, cfg.D, cfg.fH, cfg.fW)


def _gather_setup(ctx, inputs, output):
    frustum, mats, cfg_id, has_bda = inputs
    ctx.save_for_backward(mats, output[1])
    ctx.cfg_id, ctx.has_bda = cfg_id, has_bda


def _gather_backward(ctx, gout, gcnt):
    mats, cnt = ctx.saved_tensors
    return gather_pool_bwd(gout, mats, cnt, ctx.cfg_id, ctx.has_bda), None, None, None


torch.library.register_autograd("vampire_b200::gather_pool_fwd", _gather_backward, setup_context=_gather_setup)


# =============================================================================================
# render
# =============================================================================================
def _render_out_shapes(cfg: PathConfig, B: int):
    N, K, Cc = cfg.num_cams, cfg.K, cfg.C
    return [(B, N, 3, cfg.fH, cfg.fW), (B, N, K, cfg.fH, cfg.fW), (B, N, 1, cfg.fH, cfg.fW),
            (B, 3, cfg.oY, cfg.oX), (B, K, cfg.oY, cfg.oX), (B, 1, cfg.oY, cfg.oX),
            (B, 1, cfg.oZ, cfg.oY, cfg.oX), (B, Cc, cfg.oZ, cfg.oY, cfg.oX)]


def _render_in_struct(density, sem, rgb, feat, beta, geom, plan=None, flags=0, packed=None):
    rin = cabi.VbRenderIn()
    rin.flags = flags
    rin.packed = None if packed is None else packed.data_ptr()
    rin.density, rin.sem, rin.rgb, rin.feat = density.data_ptr(), sem.data_ptr(), rgb.data_ptr(), feat.data_ptr()
    rin.beta = beta.data_ptr()
    rin.geom = None if geom is None else geom.data_ptr()
    rin.plans = None if plan is None else plan.data_ptr()
    return rin


def _render_out_struct(outs):
    ro = cabi.VbRenderOut()
    for name, t in zip(("rgb", "seg", "depth", "bev_rgb", "bev_seg", "bev_height", "voxel_density", "voxel_output"),
                       outs):
        setattr(ro, name, t.data_ptr())
    return ro


def _check_render_inputs(cfg, density, sem, rgb, feat, beta):
    B = density.shape[0]
    vol = (cfg.vZ, cfg.vY, cfg.vX)
    exp = {"density_feature": (density, 1), "semantic_logits": (sem, cfg.K), "rgb": (rgb, 3),
           "voxel_features": (feat, cfg.C)}
    for name, (t, ch) in exp.items():
        if t.shape != (B, ch) + vol:
            raise ValueError(f"render: {name} has shape {tuple(t.shape)}, expected {(B, ch) + vol}")
        if t.dtype != density.dtype:
            raise TypeError("render: the four volumes must share a dtype")
    if beta.numel() != 1:
        raise ValueError("render: beta must be a scalar tensor")
    return B


def render_fwd(density: Tensor, sem: Tensor, rgb: Tensor, feat: Tensor, beta: Tensor, mats: Tensor,
               geom: Optional[Tensor], cfg_id: int, has_bda: bool, branches: int,
               plan: Optional[Tensor] = None, tanh_epilogue: bool = False) -> List[Tensor]:
    """``plan``: the device table of the batch's cached render plans (``PlanCache.render``): the camera march reads
    its sample geometry from them instead of recomputing it (forward only; the backward recomputes).
    ``tanh_epilogue``: ``voxel_output`` comes back already multiplied by ``tanh(voxel_density)`` (BV2:627-630 fused
    into the BEV kernel); inference only -- no autograd formula exists for the fused form."""
    if tanh_epilogue:
        if torch.is_grad_enabled() and any(t.requires_grad for t in (density, sem, rgb, feat, beta)):
            raise RuntimeError("render_fwd(tanh_epilogue=True) is forward-only; multiply outside when gradients are needed")
        return _render_fwd(density, sem, rgb, feat, beta, mats, geom, cfg_id, has_bda, branches | _TANH_BIT, plan)[:8]
    # the op's ninth output is its workspace: autograd keeps it (the backward re-uses the packed volume in it), every
    # other caller drops it here and the allocator gets it back at once
    return _render_fwd(density, sem, rgb, feat, beta, mats, geom, cfg_id, has_bda, branches, plan)[:8]


_TANH_BIT = 1 << 8     # rides in `branches` through the op schema (bit 0 / 1: camera / BEV branch)


@torch.library.custom_op("vampire_b200::render_fwd", mutates_args=())
def _render_fwd(density: Tensor, sem: Tensor, rgb: Tensor, feat: Tensor, beta: Tensor, mats: Tensor,
                geom: Optional[Tensor], cfg_id: int, has_bda: bool, branches: int,
                plan: Optional[Tensor]) -> List[Tensor]:
    st = state(cfg_id)
    cfg = st.cfg
    dev = _need_cuda(density, sem, rgb, feat, beta, mats, geom, plan)
    B = _check_render_inputs(cfg, density, sem, rgb, feat, beta)
    plan = _plan_ok(plan, B, dev) if geom is None else None
    dt = cabi.dtype_code(density.dtype)
    mats = _mats_ok(mats, B, cfg.num_cams)
    density, sem, rgb, feat = (t.contiguous() for t in (density, sem, rgb, feat))
    beta32 = beta.detach().reshape(1).float().contiguous()
    if geom is not None:
        if geom.shape != (B, cfg.num_cams, cfg.D, cfg.fH, cfg.fW, 3):
            raise ValueError(f"render: geom_xyz has shape {tuple(geom.shape)}")
        geom = geom.float().contiguous()
    shapes = _render_out_shapes(cfg, B)
    outs = [torch.empty(s, dtype=torch.float32, device=dev) for s in shapes[:7]]
    outs.append(torch.empty(shapes[7], dtype=density.dtype, device=dev))
    g = st.grid(B, has_bda)
    lib = cabi.lib()
    ws_bytes = _render_ws_bytes(st, g, dt, B)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    flags = cabi.RENDER_TANH_EPILOGUE if branches & _TANH_BIT else 0
    branches &= 3
    rin = _render_in_struct(density, sem, rgb, feat, beta32, geom, plan, flags)
    ro = _render_out_struct(outs)
    with torch.cuda.device(dev):
        cabi.check(lib.vb200_render_fwd(C.byref(g), C.byref(st.tables(dev).struct), mats.data_ptr(), C.byref(rin), dt,
                                        C.byref(ro), branches, ws.data_ptr(), ws_bytes, cabi.stream_ptr(dev)))
    if not branches & cabi.BRANCH_CAM:
        for t in outs[:3]:
            t.zero_()
    if not branches & cabi.BRANCH_BEV:
        for t in outs[3:]:
            t.zero_()
    return outs + [ws]


def _render_ws_bytes(st, g, dt, B):
    lib = cabi.lib()
    return lib.vb200_render_fwd_workspace(C.byref(g), dt) + \
        lib.vb200_render_packed_bytes(C.byref(g), dt) * ((B if st.render_group <= 0 else min(B, st.render_group)) - 1)


@_render_fwd.register_fake
def _(density, sem, rgb, feat, beta, mats, geom, cfg_id, has_bda, branches, plan):
    st = state(cfg_id)
    cfg = st.cfg
    B = density.shape[0]
    shapes = _render_out_shapes(cfg, B)
    outs = [density.new_empty(s, dtype=torch.float32) for s in shapes[:7]]
    outs.append(density.new_empty(shapes[7]))
    ws_bytes = _render_ws_bytes(st, st.grid(B, has_bda), cabi.dtype_code(density.dtype), B)
    return outs + [density.new_empty(ws_bytes, dtype=torch.uint8)]


@torch.library.custom_op("vampire_b200::render_bwd", mutates_args=())
def render_bwd(grads: List[Tensor], outs: List[Tensor], density: Tensor, sem: Tensor, rgb: Tensor, feat: Tensor,
               beta: Tensor, mats: Tensor, geom: Optional[Tensor], cfg_id: int, has_bda: bool,
               branches: int, packed: Optional[Tensor]) -> List[Tensor]:
    """``packed``: the forward's workspace when it holds the channels-last copy of all B samples (else None): the
    backward then reads that copy instead of packing density | sem | rgb a second time."""
    st = state(cfg_id)
    cfg = st.cfg
    dev = _need_cuda(density, sem, rgb, feat, beta, mats, geom)
    B = density.shape[0]
    dt = cabi.dtype_code(density.dtype)
    mats = _mats_ok(mats, B, cfg.num_cams)
    density, sem, rgb, feat = (t.contiguous() for t in (density, sem, rgb, feat))
    beta32 = beta.detach().reshape(1).float().contiguous()
    if geom is not None:
        geom = geom.float().contiguous()
    grads = [gt.float().contiguous() for gt in grads[:7]] + [grads[7].to(density.dtype).contiguous()]
    outs = [o.contiguous() for o in outs]
    g_den, g_sem, g_rgb, g_feat = (torch.empty_like(t) for t in (density, sem, rgb, feat))
    g_beta = torch.zeros(1, dtype=torch.float32, device=dev)
    g = st.grid(B, has_bda)
    lib = cabi.lib()
    ws_bytes = lib.vb200_render_bwd_workspace(C.byref(g), dt)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    if packed is not None:
        need = lib.vb200_render_packed_bytes(C.byref(g), dt) * B
        if packed.dtype != torch.uint8 or packed.numel() < need or packed.device != dev:
            raise ValueError("render_bwd: `packed` is not the forward's workspace of this batch")
    rin = _render_in_struct(density, sem, rgb, feat, beta32, geom, packed=packed)
    ro = _render_out_struct(outs)
    rg = cabi.VbRenderGrad()
    for name, t in zip(("g_rgb", "g_seg", "g_depth", "g_bev_rgb", "g_bev_seg", "g_bev_height", "g_voxel_density",
                        "g_voxel_output"), grads):
        setattr(rg, name, t.data_ptr())
    rg.g_density, rg.g_sem, rg.g_rgb_in, rg.g_feat = (t.data_ptr() for t in (g_den, g_sem, g_rgb, g_feat))
    rg.g_beta = g_beta.data_ptr()
    with torch.cuda.device(dev):
        cabi.check(lib.vb200_render_bwd(C.byref(g), C.byref(st.tables(dev).struct), mats.data_ptr(), C.byref(rin), dt,
                                        C.byref(ro), C.byref(rg), branches, ws.data_ptr(), ws_bytes,
                                        cabi.stream_ptr(dev)))
    return [g_den, g_sem, g_rgb, g_feat, g_beta]


@render_bwd.register_fake
def _(grads, outs, density, sem, rgb, feat, beta, mats, geom, cfg_id, has_bda, branches, packed):
    return [torch.empty_like(density), torch.empty_like(sem), torch.empty_like(rgb), torch.empty_like(feat),
            density.new_empty(1, dtype=torch.float32)]


def _render_setup(ctx, inputs, output):
    density, sem, rgb, feat, beta, mats, geom, cfg_id, has_bda, branches, _plan = inputs
    ctx.save_for_backward(density, sem, rgb, feat, beta, mats, geom, *output)
    ctx.cfg_id, ctx.has_bda, ctx.branches = cfg_id, has_bda, branches
    # the workspace (last output) starts with the packed copies of ALL samples only when one pack round covered the batch
    st = state(cfg_id)
    ctx.packed_ok = bool(branches & cabi.BRANCH_CAM) and (st.render_group <= 0 or st.render_group >= density.shape[0])


def _render_backward(ctx, grads):
    saved = ctx.saved_tensors
    density, sem, rgb, feat, beta, mats, geom = saved[:7]
    outs, ws = list(saved[7:15]), saved[15]
    grads = [gt if gt is not None else torch.zeros_like(o) for gt, o in zip(grads[:8], outs)]
    g_den, g_sem, g_rgb, g_feat, g_beta = render_bwd(grads, outs, density, sem, rgb, feat, beta, mats, geom,
                                                     ctx.cfg_id, ctx.has_bda, ctx.branches & 3,
                                                     ws if ctx.packed_ok else None)
    return g_den, g_sem, g_rgb, g_feat, g_beta.reshape(beta.shape).to(beta.dtype), None, None, None, None, None, None


torch.library.register_autograd("vampire_b200::render_fwd", _render_backward, setup_context=_render_setup)


# =============================================================================================
# producer right before the path: softmax over the depth planes, BV2:551
@torch.library.custom_op("vampire_b200::depth_softmax_fwd", mutates_args=())
def depth_softmax_fwd(logits: Tensor, out_fp32: bool) -> Tensor:
    """logits (..., D, fH, fW) -> softmax over D (dim=-3), fp32 arithmetic; output in fp32 or the logits' dtype."""
    dev = _need_cuda(logits)
    if logits.dim() < 3:
        raise ValueError("depth_softmax: need (..., D, fH, fW)")
    x = logits.contiguous()
    D, inner = x.shape[-3], x.shape[-2] * x.shape[-1]
    outer = x.numel() // (D * inner) if x.numel() else 0
    out = torch.empty_like(x, dtype=torch.float32 if out_fp32 else x.dtype)
    if x.numel():
        with torch.cuda.device(dev):
            cabi.check(cabi.lib().vb200_depth_softmax_fwd(
                x.data_ptr(), cabi.dtype_code(x.dtype), out.data_ptr(), cabi.dtype_code(out.dtype), outer, D, inner,
                cabi.stream_ptr(dev)))
    return out


@depth_softmax_fwd.register_fake
def _(logits, out_fp32):
    return torch.empty_like(logits, dtype=torch.float32 if out_fp32 else logits.dtype,
                            memory_format=torch.contiguous_format)


@torch.library.custom_op("vampire_b200::depth_softmax_bwd", mutates_args=())
def depth_softmax_bwd(probs: Tensor, gprobs: Tensor, out_dtype: torch.dtype) -> Tensor:
    dev = _need_cuda(probs, gprobs)
    y = probs.contiguous()
    g = gprobs.to(y.dtype).contiguous()
    D, inner = y.shape[-3], y.shape[-2] * y.shape[-1]
    outer = y.numel() // (D * inner) if y.numel() else 0
    if y.dtype != torch.float32 and out_dtype != y.dtype:
        raise TypeError("depth_softmax_bwd: a 16-bit probability tensor yields gradients of the same dtype")
    out = torch.empty_like(y, dtype=out_dtype)
    if y.numel():
        with torch.cuda.device(dev):
            cabi.check(cabi.lib().vb200_depth_softmax_bwd(
                y.data_ptr(), g.data_ptr(), cabi.dtype_code(y.dtype), out.data_ptr(), cabi.dtype_code(out_dtype),
                outer, D, inner, cabi.stream_ptr(dev)))
    return out


@depth_softmax_bwd.register_fake
def _(probs, gprobs, out_dtype):
    return torch.empty_like(probs, dtype=out_dtype, memory_format=torch.contiguous_format)


def _smx_setup(ctx, inputs, output):
    ctx.in_dtype = inputs[0].dtype
    ctx.save_for_backward(output)


def _smx_backward(ctx, gout):
    (y,) = ctx.saved_tensors
    return depth_softmax_bwd(y, gout, ctx.in_dtype), None


torch.library.register_autograd("vampire_b200::depth_softmax_fwd", _smx_backward, setup_context=_smx_setup)


# callers right after the path: x4 upsample of the rendered maps, point queries
# =============================================================================================
@torch.library.custom_op("vampire_b200::upsample_fwd", mutates_args=())
def upsample_fwd(x: Tensor, factor: int) -> Tensor:
    dev = _need_cuda(x)
    if x.dim() < 2:
        raise ValueError("upsample: need (..., H, W)")
    H, W = x.shape[-2:]
    xin = x.float().contiguous()
    planes = xin.numel() // (H * W)
    out = torch.empty(x.shape[:-2] + (H * factor, W * factor), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        for p0 in range(0, planes, 65535):
            n = min(65535, planes - p0)
            cabi.check(cabi.lib().vb200_upsample_bilinear_fwd(
                xin.data_ptr() + p0 * H * W * 4, out.data_ptr() + p0 * H * W * factor * factor * 4, n, H, W, factor,
                cabi.stream_ptr(dev)))
    return out


@upsample_fwd.register_fake
def _(x, factor):
    return x.new_empty(x.shape[:-2] + (x.shape[-2] * factor, x.shape[-1] * factor), dtype=torch.float32)


@torch.library.custom_op("vampire_b200::upsample_bwd", mutates_args=())
def upsample_bwd(gout: Tensor, factor: int) -> Tensor:
    dev = _need_cuda(gout)
    OH, OW = gout.shape[-2:]
    H, W = OH // factor, OW // factor
    g = gout.float().contiguous()
    planes = g.numel() // (OH * OW)
    gin = torch.empty(gout.shape[:-2] + (H, W), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        for p0 in range(0, planes, 65535):
            n = min(65535, planes - p0)
            cabi.check(cabi.lib().vb200_upsample_bilinear_bwd(
                g.data_ptr() + p0 * OH * OW * 4, gin.data_ptr() + p0 * H * W * 4, n, H, W, factor,
                cabi.stream_ptr(dev)))
    return gin


@upsample_bwd.register_fake
def _(gout, factor):
    return gout.new_empty(gout.shape[:-2] + (gout.shape[-2] // factor, gout.shape[-1] // factor), dtype=torch.float32)


def _up_setup(ctx, inputs, output):
    ctx.factor = inputs[1]
    ctx.in_dtype = inputs[0].dtype


def _up_backward(ctx, gout):
    return upsample_bwd(gout, ctx.factor).to(ctx.in_dtype), None


torch.library.register_autograd("vampire_b200::upsample_fwd", _up_backward, setup_context=_up_setup)


@torch.library.custom_op("vampire_b200::query_points_fwd", mutates_args=())
def query_points_fwd(vol: Tensor, pts: Tensor, rot: Optional[Tensor], beta: Optional[Tensor], cfg_id: int,
                     border: bool, apply_density: bool, mask_invalid: bool) -> Tuple[Tensor, Tensor]:
    """vol (B,CH,Z,Y,X); pts (B,P,3) or (P,3) ego coordinates -> (out (B,CH,P) fp32, valid (B,P) uint8)."""
    st = state(cfg_id)
    cfg = st.cfg
    dev = _need_cuda(vol, pts, rot, beta)
    B, CH = vol.shape[:2]
    if vol.shape[2:] != (cfg.vZ, cfg.vY, cfg.vX):
        raise ValueError(f"query_points: volume {tuple(vol.shape)} does not match the seg grid")
    batched = pts.dim() == 3
    P = pts.shape[-2]
    pts32 = pts.float().contiguous()
    vol = vol.contiguous()
    rot32 = None if rot is None else rot.float().reshape(B, 9).contiguous()
    beta32 = None if beta is None else beta.detach().reshape(1).float().contiguous()
    out = torch.empty(B, CH, P, dtype=torch.float32, device=dev)
    valid = torch.empty(B, P, dtype=torch.uint8, device=dev)
    g = st.grid(B, True)
    with torch.cuda.device(dev):
        cabi.check(cabi.lib().vb200_query_points_fwd(
            C.byref(g), vol.data_ptr(), cabi.dtype_code(vol.dtype), CH, pts32.data_ptr(), P, int(batched),
            cabi.ptr(rot32), int(border), int(apply_density), int(mask_invalid), cabi.ptr(beta32), out.data_ptr(),
            valid.data_ptr(), cabi.stream_ptr(dev)))
    return out, valid


@query_points_fwd.register_fake
def _(vol, pts, rot, beta, cfg_id, border, apply_density, mask_invalid):
    B, CH = vol.shape[:2]
    P = pts.shape[-2]
    return vol.new_empty(B, CH, P, dtype=torch.float32), vol.new_empty(B, P, dtype=torch.uint8)


@torch.library.custom_op("vampire_b200::query_points_bwd", mutates_args=())
def query_points_bwd(gout: Tensor, vol: Tensor, pts: Tensor, rot: Optional[Tensor], beta: Optional[Tensor],
                     cfg_id: int, border: bool, apply_density: bool, mask_invalid: bool) -> Tuple[Tensor, Tensor]:
    st = state(cfg_id)
    dev = _need_cuda(gout, vol, pts, rot, beta)
    B, CH = vol.shape[:2]
    batched = pts.dim() == 3
    P = pts.shape[-2]
    pts32 = pts.float().contiguous()
    vol = vol.contiguous()
    gout = gout.float().contiguous()
    rot32 = None if rot is None else rot.float().reshape(B, 9).contiguous()
    beta32 = None if beta is None else beta.detach().reshape(1).float().contiguous()
    gvol = torch.empty_like(vol)
    gbeta = torch.zeros(1, dtype=torch.float32, device=dev)
    g = st.grid(B, True)
    lib = cabi.lib()
    dt = cabi.dtype_code(vol.dtype)
    ws_bytes = lib.vb200_query_points_bwd_workspace(C.byref(g), CH, P, dt)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        cabi.check(lib.vb200_query_points_bwd(
            C.byref(g), vol.data_ptr(), dt, CH, pts32.data_ptr(), P, int(batched), cabi.ptr(rot32), int(border),
            int(apply_density), int(mask_invalid), cabi.ptr(beta32), gout.data_ptr(), gvol.data_ptr(), gbeta.data_ptr(),
            ws.data_ptr(), ws_bytes, cabi.stream_ptr(dev)))
    return gvol, gbeta


@query_points_bwd.register_fake
def _(gout, vol, pts, rot, beta, cfg_id, border, apply_density, mask_invalid):
    return torch.empty_like(vol), vol.new_empty(1, dtype=torch.float32)


def _query_setup(ctx, inputs, output):
    vol, pts, rot, beta, cfg_id, border, apply_density, mask_invalid = inputs
    ctx.save_for_backward(vol, pts, rot, beta)
    ctx.args = (cfg_id, border, apply_density, mask_invalid)


def _query_backward(ctx, gout, gvalid):
    vol, pts, rot, beta = ctx.saved_tensors
    gvol, gbeta = query_points_bwd(gout, vol, pts, rot, beta, *ctx.args)
    gb = None if beta is None else gbeta.reshape(beta.shape).to(beta.dtype)
    return gvol, None, None, gb, None, None, None, None


torch.library.register_autograd("vampire_b200::query_points_fwd", _query_backward, setup_context=_query_setup)
