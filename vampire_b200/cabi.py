"""ctypes binding of libvb200.so -- the C ABI declared in ``include/vb200.h``.

This is the only place Python touches the native library.  There is deliberately no fallback:
if the library is missing or the device is not a B200 every compute call raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

from .config import PathConfig
from .lattice import Lattice

_HERE = os.path.dirname(os.path.abspath(__file__))
# VB200_LIB: load an alternative build of the same library (kernel-tuning experiments: tools/build_variant.sh)
LIBPATH = os.environ.get("VB200_LIB") or os.path.join(_HERE, "_lib", "libvb200.so")

F32, BF16, F16 = 0, 1, 2
NCDHW, NDHWC = 0, 1
BRANCH_CAM, BRANCH_BEV = 1, 2

_DTYPES = {torch.float32: F32, torch.bfloat16: BF16, torch.float16: F16}


class VbGrid(C.Structure):
    _fields_ = [
        ("B", C.c_int32), ("N", C.c_int32),
        ("D", C.c_int32), ("fH", C.c_int32), ("fW", C.c_int32),
        ("vZ", C.c_int32), ("vY", C.c_int32), ("vX", C.c_int32),
        ("oZ", C.c_int32), ("oY", C.c_int32), ("oX", C.c_int32),
        ("C", C.c_int32), ("K", C.c_int32), ("has_bda", C.c_int16), ("density_mode", C.c_int16),
        ("img_w_m1", C.c_float), ("img_h_m1", C.c_float),
        ("x_hi", C.c_float), ("y_hi", C.c_float),
        ("d_lo", C.c_float), ("d_hi", C.c_float), ("d_ext", C.c_float),
        ("seg_lo", C.c_float * 3), ("seg_ext", C.c_float * 3),
        ("bg_depth", C.c_float), ("bev_delta", C.c_float),
        ("sdf_bias", C.c_float), ("beta_min", C.c_float), ("term_eps", C.c_float),
    ]


class VbTables(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in
                ("us", "vs", "ds", "xs", "ys", "zs", "oxs", "oys", "ozs", "mids", "bev_mids")]


class VbLiftPlan(C.Structure):
    """One sample's cached lift plan (device pointers); see include/vb200.h."""
    _fields_ = [(n, C.c_void_p) for n in ("head", "pairs", "cell_off", "cell_recs")]


class VbRenderPlan(C.Structure):
    """One sample's cached render plan (device pointers); see include/vb200.h."""
    _fields_ = [(n, C.c_void_p) for n in ("steps", "delta", "last", "box")]


class VbRenderIn(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("density", "sem", "rgb", "feat", "beta", "geom", "plans")] + \
        [("flags", C.c_int32), ("packed", C.c_void_p)]


RENDER_TANH_EPILOGUE = 1


class VbRenderOut(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in
                ("rgb", "seg", "depth", "bev_rgb", "bev_seg", "bev_height", "voxel_density", "voxel_output")]


class VbRenderGrad(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in
                ("g_rgb", "g_seg", "g_depth", "g_bev_rgb", "g_bev_seg", "g_bev_height", "g_voxel_density",
                 "g_voxel_output", "g_density", "g_sem", "g_rgb_in", "g_feat", "g_beta")]


_P = C.c_void_p
_PROTOS = {
    "vb200_version": (C.c_int, []),
    "vb200_strerror": (C.c_char_p, [C.c_int]),
    "vb200_device_check": (C.c_int, []),
    "vb200_trace_num_kernels": (C.c_int, []),
    "vb200_trace_kernel_name": (C.c_char_p, [C.c_int]),
    "vb200_launch_count": (C.c_longlong, [C.c_int]),
    "vb200_trace_enable": (C.c_int, [C.c_int]),
    "vb200_trace_collect": (C.c_int, [C.POINTER(C.c_double), C.POINTER(C.c_longlong)]),
    "vb200_get_pixel": (C.c_int, [C.POINTER(VbGrid), C.POINTER(VbTables), _P, _P, _P]),
    "vb200_get_geometry": (C.c_int, [C.POINTER(VbGrid), C.POINTER(VbTables), _P, _P, C.c_int, _P]),
    "vb200_lift_indices": (C.c_int, [C.POINTER(VbGrid), C.POINTER(VbTables), _P, _P, _P, _P, _P]),
    "vb200_render_indices": (C.c_int, [C.POINTER(VbGrid), C.POINTER(VbTables), _P, _P, _P, _P, _P, _P]),
    "vb200_lift_pool_fwd_workspace": (C.c_size_t, [C.POINTER(VbGrid), C.c_int]),
    "vb200_lift_pool_bwd_workspace": (C.c_size_t, [C.POINTER(VbGrid), C.c_int]),
    "vb200_lift_pool_fwd": (C.c_int, [C.POINTER(VbGrid), C.POINTER(VbTables), _P, _P, _P, C.c_int, C.c_int, _P,
                                      C.c_int, _P, _P, C.c_size_t, _P]),
    "vb200_lift_pool_bwd": (C.c_int, [C.POINTER(VbGrid), C.POINTER(VbTables), _P, _P, _P, C.c_int, C.c_int, _P,
                                      C.c_int, _P, _P, _P, _P, C.c_size_t, _P]),
    "vb200_lift_plan_workspace": (C.c_size_t, [C.POINTER(VbGrid), C.c_longlong]),
    "vb200_lift_plan_build": (C.c_int, [C.POINTER(VbGrid), C.POINTER(VbTables), _P, _P, _P, _P, _P, C.c_longlong, _P,
                                        _P, C.c_size_t, _P]),
    "vb200_lift_pool_bwd_planned_workspace": (C.c_size_t, [C.POINTER(VbGrid), C.c_int]),
    "vb200_lift_pool_fwd_planned": (C.c_int, [C.POINTER(VbGrid), _P, _P, _P, C.c_int, C.c_int, _P, C.c_int, _P, _P,
                                              C.c_size_t, _P]),
    "vb200_lift_pool_bwd_planned": (C.c_int, [C.POINTER(VbGrid), _P, _P, _P, C.c_int, C.c_int, _P, C.c_int, _P, _P,
                                              _P, _P, C.c_size_t, _P]),
    "vb200_gather_pool_bwd_workspace": (C.c_size_t, [C.POINTER(VbGrid), C.c_int]),
    "vb200_gather_pool_fwd": (C.c_int, [C.POINTER(VbGrid), C.POINTER(VbTables), _P, _P, C.c_int, _P, _P, _P]),
    "vb200_gather_pool_bwd": (C.c_int, [C.POINTER(VbGrid), C.POINTER(VbTables), _P, _P, _P, C.c_int, _P, _P,
                                        C.c_size_t, _P]),
    "vb200_depth_softmax_fwd": (C.c_int, [_P, C.c_int, _P, C.c_int, C.c_longlong, C.c_int, C.c_int, _P]),
    "vb200_depth_softmax_bwd": (C.c_int, [_P, _P, C.c_int, _P, C.c_int, C.c_longlong, C.c_int, C.c_int, _P]),
    "vb200_upsample_bilinear_fwd": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    "vb200_upsample_bilinear_bwd": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    "vb200_query_points_bwd_workspace": (C.c_size_t, [C.POINTER(VbGrid), C.c_int, C.c_int, C.c_int]),
    "vb200_query_points_fwd": (C.c_int, [C.POINTER(VbGrid), _P, C.c_int, C.c_int, _P, C.c_int, C.c_int, _P, C.c_int,
                                         C.c_int, C.c_int, _P, _P, _P, _P]),
    "vb200_query_points_bwd": (C.c_int, [C.POINTER(VbGrid), _P, C.c_int, C.c_int, _P, C.c_int, C.c_int, _P, C.c_int,
                                         C.c_int, C.c_int, _P, _P, _P, _P, _P, C.c_size_t, _P]),
    "vb200_render_plan_rays": (C.c_size_t, [C.POINTER(VbGrid)]),
    "vb200_render_plan_build": (C.c_int, [C.POINTER(VbGrid), C.POINTER(VbTables), _P, _P, _P, _P, _P, _P]),
    "vb200_render_fwd_workspace": (C.c_size_t, [C.POINTER(VbGrid), C.c_int]),
    "vb200_render_bwd_workspace": (C.c_size_t, [C.POINTER(VbGrid), C.c_int]),
    "vb200_render_packed_bytes": (C.c_size_t, [C.POINTER(VbGrid), C.c_int]),
    "vb200_render_set_fork": (C.c_int, [C.c_int]),
    "vb200_render_set_march_split": (C.c_int, [C.c_int]),
    "vb200_render_fwd": (C.c_int, [C.POINTER(VbGrid), C.POINTER(VbTables), _P, C.POINTER(VbRenderIn), C.c_int,
                                   C.POINTER(VbRenderOut), C.c_int, _P, C.c_size_t, _P]),
    "vb200_render_bwd": (C.c_int, [C.POINTER(VbGrid), C.POINTER(VbTables), _P, C.POINTER(VbRenderIn), C.c_int,
                                   C.POINTER(VbRenderOut), C.POINTER(VbRenderGrad), C.c_int, _P, C.c_size_t, _P]),
}

_lib: Optional[C.CDLL] = None


def exported_symbols():
    """Every entry point ``include/vb200.h`` declares."""
    return sorted(_PROTOS.keys())


def lib() -> C.CDLL:
    """Load libvb200.so (built in-tree by ``python -m vampire_b200.build``).  Fails loudly."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIBPATH):
            raise RuntimeError(
                f"libvb200.so not found at {LIBPATH}: build it with `python -m vampire_b200.build`. "
                "vampire_b200 has no CPU or PyTorch fallback for the lift/pool/render path.")
        handle = C.CDLL(LIBPATH)
        for name, (res, args) in _PROTOS.items():
            fn = getattr(handle, name)   # AttributeError if the header and the library disagree
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        raise RuntimeError(lib().vb200_strerror(rc).decode())


def dtype_code(dt: torch.dtype) -> int:
    try:
        return _DTYPES[dt]
    except KeyError:
        raise TypeError(f"vampire_b200: unsupported feature dtype {dt}") from None


DENSITY_MODES = {"sdf": 0, "naive": 1}       # enum vb200_density


def make_grid(cfg: PathConfig, batch: int, has_bda: bool = True, term_eps: float = 1e-8,
              lift_2d: bool = False) -> VbGrid:
    """Sizes + fp32 constants; each float is the reference's Python double, rounded by c_float
    exactly as torch rounds a Python scalar operand of an fp32 tensor op."""
    H, W = cfg.final_dim
    g = VbGrid()
    g.B, g.N = batch, cfg.num_cams
    g.D, g.fH, g.fW = cfg.D, cfg.fH, cfg.fW
    g.vZ, g.vY, g.vX = cfg.vZ, cfg.vY, cfg.vX
    g.oZ, g.oY, g.oX = cfg.oZ, cfg.oY, cfg.oX
    g.C, g.K = cfg.C, cfg.K
    g.has_bda = 1 if has_bda else 0
    g.img_w_m1, g.img_h_m1 = float(W - 1), float(H - 1)          # BV2:499-500
    g.x_hi, g.y_hi = float(W - 0.5), float(H - 0.5)              # BV2:494-495
    g.d_lo, g.d_hi = cfg.d_bound[0], cfg.d_bound[1]              # BV2:496
    g.d_ext = cfg.d_bound[1] - cfg.d_bound[0]                    # BV2:501 (double subtraction, then fp32)
    lo = (cfg.x_bound_seg[0], cfg.y_bound_seg[0], cfg.z_bound_seg[0])
    ext = (cfg.x_bound_seg[1] - cfg.x_bound_seg[0], cfg.y_bound_seg[1] - cfg.y_bound_seg[0],
           cfg.z_bound_seg[1] - cfg.z_bound_seg[0])              # BV2:397-402
    for i in range(3):
        g.seg_lo[i] = lo[i]
        g.seg_ext[i] = ext[i]
    g.bg_depth = cfg.d_bound[1]                                  # BV2:436
    g.bev_delta = cfg.z_bound_det[2]                             # BV2:451
    g.sdf_bias = cfg.sdf_bias
    g.beta_min = 1e-4                                            # render_utils.py:31
    g.term_eps = term_eps
    g.density_mode = DENSITY_MODES[cfg.density_mode]             # BV2:191-194
    if lift_2d:
        # BaseBiLinear's 2-D lift (base_bilinear.py:471-517): D == 1 selects it -- one depth plane, test z > 0
        g.D = 1
        g.d_lo, g.d_hi, g.d_ext = 0.0, float("inf"), 1.0
    return g


class DeviceTables:
    """The lattice tables resident on one CUDA device + the VbTables struct pointing at them."""

    def __init__(self, lat: Lattice, device: torch.device):
        self.packed = lat.packed().to(device)
        self.struct = VbTables()
        off = 0
        base = self.packed.data_ptr()
        for name in ("us", "vs", "ds", "xs", "ys", "zs", "oxs", "oys", "ozs", "mids", "bev_mids"):
            setattr(self.struct, name, base + 4 * off)
            off += getattr(lat, name).numel()


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def launch_count() -> int:
    """Kernels launched by libvb200 since it was loaded."""
    return int(lib().vb200_launch_count(-1))


def trace_enable(on: bool) -> None:
    check(lib().vb200_trace_enable(1 if on else 0))


def trace_collect():
    """{kernel family: (device ms total, launches)} for the launches recorded since trace_enable."""
    l = lib()
    n = l.vb200_trace_num_kernels()
    ms = (C.c_double * n)()
    cnt = (C.c_longlong * n)()
    check(l.vb200_trace_collect(ms, cnt))
    return {l.vb200_trace_kernel_name(i).decode(): (ms[i], int(cnt[i])) for i in range(n) if cnt[i]}


def render_set_march_split(segments: int) -> None:
    """Depth segments of the camera march: 0 = default (never), 1 = never, -1 = automatic, 2 / 4 / 8 = forced
    (include/vb200.h)."""
    check(lib().vb200_render_set_march_split(int(segments)))


def render_set_fork(enable: bool) -> None:
    """Fork the BEV branch onto libvb200's side stream (default) or serialise the two branches."""
    check(lib().vb200_render_set_fork(1 if enable else 0))
