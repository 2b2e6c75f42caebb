"""ORACLE (test infrastructure only): NumPy fixed-order fp32 restatement.

NumPy evaluates each fp32 ufunc with one IEEE rounding and never contracts a*b+c into an FMA,
so writing the reference's arithmetic out operation by operation reproduces the CPU reference
bit for bit while exposing the integers ``F.grid_sample`` hides: corner base
indices, in-bounds bits, validity masks, fractions.  The same
operation order is what the CUDA kernels implement with ``__fmul_rn/__fadd_rn/__fdiv_rn``.

Inputs are the *prepared* matrices (vampire_b200.matrices layout, (B,N,6,4,4)); the 4x4
inverse/products themselves are the reference's own torch calls (BV2:334,340,374,379) and are
not restated -- LAPACK's LU is a third-party dependency of the reference (torch), called
identically on both sides.
"""
from __future__ import annotations

import numpy as np

F32 = np.float32


def _mv(M, p):
    """Row-major 4x4 times 4-vector, ATen native bmm order: acc = 0; acc += M[i,k]*p[k], k=0..3
    with separate multiply and add roundings.  M: (...,4,4) broadcastable
    against the leading dims of the components p[k]."""
    out = []
    for i in range(4):
        acc = M[..., i, 0] * p[0]
        acc = acc + M[..., i, 1] * p[1]
        acc = acc + M[..., i, 2] * p[2]
        acc = acc + M[..., i, 3] * p[3]
        out.append(acc.astype(F32, copy=False))
    return out


def project_voxels(prep, xs, ys, zs, has_bda=True):
    """G1 ``get_pixel`` BV2:351-388.  prep (B,N,6,4,4) fp32 -> (B,N,Z,Y,X,3) fp32."""
    prep = np.asarray(prep, dtype=F32)
    B, N = prep.shape[:2]
    X = np.asarray(xs, F32)[None, None, None, None, :]
    Y = np.asarray(ys, F32)[None, None, None, :, None]
    Z = np.asarray(zs, F32)[None, None, :, None, None]
    one = F32(1.0)
    shape = (B, N, Z.shape[2], Y.shape[3], X.shape[4])
    p = [np.broadcast_to(X, shape), np.broadcast_to(Y, shape), np.broadcast_to(Z, shape),
         np.broadcast_to(one, shape)]
    m = lambda s: prep[:, :, s][:, :, None, None, None]
    if has_bda:
        p = _mv(m(0), p)                       # BV2:374
    p = _mv(m(1), p)                           # BV2:380
    zc = np.maximum(p[2], F32(1e-6))           # BV2:385 clamp(min=eps)
    q = [p[0] / zc, p[1] / zc, p[2], p[3]]
    r = _mv(m(2), q)                           # BV2:387
    return np.stack(r[:3], axis=-1)


def frustum_points(prep, us, vs, ds, has_bda=True):
    """G2 ``get_geometry`` BV2:314-349.  -> (B,N,D,fH,fW,3) fp32."""
    prep = np.asarray(prep, dtype=F32)
    B, N = prep.shape[:2]
    U = np.asarray(us, F32)[None, None, None, None, :]
    V = np.asarray(vs, F32)[None, None, None, :, None]
    Dd = np.asarray(ds, F32)[None, None, :, None, None]
    shape = (B, N, Dd.shape[2], V.shape[3], U.shape[4])
    p = [np.broadcast_to(U, shape), np.broadcast_to(V, shape), np.broadcast_to(Dd, shape),
         np.broadcast_to(F32(1.0), shape)]
    m = lambda s: prep[:, :, s][:, :, None, None, None]
    p = _mv(m(3), p)                           # BV2:334
    p = [p[0] * p[2], p[1] * p[2], p[2], p[3]]  # BV2:336-338
    p = _mv(m(4), p)                           # BV2:341-342
    if has_bda:
        p = _mv(m(5), p)                       # BV2:346
    return np.stack(p[:3], axis=-1)


def lift_indices(pix, final_dim, d_bound, sizes):
    """L2 (BV2:493-506) + ATen unnormalise, align_corners=False.

    pix (...,3) fp32; sizes=(W,H,D) of the sampled frustum volume.
    Returns dict: valid(bool), n(3 fp32), i(3 fp32 unnormalised), i0(3 int32), f(3 fp32).
    """
    H, W = final_dim
    x, y, z = pix[..., 0], pix[..., 1], pix[..., 2]
    valid = (x > F32(-0.5)) & (x < F32(W - 0.5)) & (y > F32(-0.5)) & (y < F32(H - 0.5)) & \
            (z > F32(d_bound[0])) & (z < F32(d_bound[1]))
    nx = F32(2.0) * (x / F32(float(W - 1))) - F32(1.0)
    ny = F32(2.0) * (y / F32(float(H - 1))) - F32(1.0)
    nz = F32(2.0) * ((z - F32(d_bound[0])) / F32(d_bound[1] - d_bound[0])) - F32(1.0)
    n = [np.clip(v, F32(-2.0), F32(2.0)) for v in (nx, ny, nz)]
    out = {"valid": valid, "n": n, "i": [], "i0": [], "f": []}
    for v, size in zip(n, sizes):
        i = ((v + F32(1.0)) * F32(size) - F32(1.0)) / F32(2.0)   # GridSampler.h unnormalize
        i0 = np.floor(i)
        out["i"].append(i)
        out["i0"].append(i0.astype(np.int32))
        out["f"].append((i - i0).astype(F32))
    return out


def render_indices(geom, lo, ext, sizes):
    """R2 (BV2:397-407) + ATen unnormalise, align_corners=True.

    geom (...,3) fp32 (already nan_to_num'ed, BV2:612); lo/ext = fp32 triples built from Python
    doubles like the reference's ``torch.as_tensor([...])``; sizes=(X,Y,Z).
    """
    out = {"g": [], "i": [], "i0": [], "f": []}
    mask = None
    for a in range(3):
        g = (geom[..., a] - F32(lo[a])) / F32(ext[a])
        g = g * F32(2.0) - F32(1.0)
        ok = (g >= F32(-1.0)) & (g <= F32(1.0))
        mask = ok if mask is None else (mask & ok)
        i = ((g + F32(1.0)) / F32(2.0)) * F32(sizes[a] - 1)
        i0 = np.floor(i)
        out["g"].append(g)
        out["i"].append(i)
        with np.errstate(invalid="ignore"):
            out["i0"].append(np.where(np.isfinite(i0), i0, 0).astype(np.int64).clip(-2**31, 2**31 - 1).astype(np.int32))
        out["f"].append((i - i0).astype(F32))
    out["mask"] = mask
    return out


def nan_to_num(a, nan=0.0):
    """torch.nan_to_num for fp32: nan -> nan arg, +-inf -> +-FLT_MAX (BV2:421, 612)."""
    fmax = np.finfo(F32).max
    return np.nan_to_num(a, nan=F32(nan), posinf=fmax, neginf=-fmax).astype(F32)


# ---- factorised lift, small sizes only --------------------------------------
def lift_pool_factorised(depth, ctx, pix, final_dim, d_bound):
    """depth (B,N,D,h,w), ctx (B,N,C,h,w), pix (B,N,Z,Y,X,3) -> (B,C,Z,Y,X), cnt (B,C,Z,Y,X)."""
    depth = np.asarray(depth, F32)
    ctx = np.asarray(ctx, F32)
    B, N, D, h, w = depth.shape
    C = ctx.shape[2]
    idx = lift_indices(pix, final_dim, d_bound, (w, h, D))
    x0, y0, z0 = idx["i0"]
    fx, fy, fz = idx["f"]
    valid = idx["valid"]
    Z, Y, X = pix.shape[2:5]
    numer = np.zeros((B, C, Z, Y, X), F32)
    cnt = np.zeros((B, C, Z, Y, X), np.int32)
    bb = np.arange(B)[:, None, None, None]
    for n in range(N):
        f = np.zeros((B, C, Z, Y, X), F32)
        for dy in (0, 1):
            for dx in (0, 1):
                yy = y0[:, n] + dy
                xx = x0[:, n] + dx
                inb = (yy >= 0) & (yy < h) & (xx >= 0) & (xx < w)
                yc = np.clip(yy, 0, h - 1)
                xc = np.clip(xx, 0, w - 1)
                wy = fy[:, n] if dy else (F32(1) - fy[:, n])
                wx = fx[:, n] if dx else (F32(1) - fx[:, n])
                s = np.zeros(yy.shape, F32)
                for dz in (0, 1):
                    zz = z0[:, n] + dz
                    inz = (zz >= 0) & (zz < D)
                    zc = np.clip(zz, 0, D - 1)
                    wz = fz[:, n] if dz else (F32(1) - fz[:, n])
                    s = s + np.where(inz, wz * depth[bb, n, zc, yc, xc], F32(0))
                wgt = np.where(inb & valid[:, n], wy * wx * s, F32(0)).astype(F32)
                cv = ctx[:, n][bb[:, None], np.arange(C)[None, :, None, None, None], yc[:, None], xc[:, None]]
                f = f + cv * wgt[:, None]
        numer += f
        cnt += (np.abs(f) > 0)
    out = numer / (cnt.astype(F32) + F32(1e-6))
    return out.astype(F32), cnt


# ---- density + sequential compositing, small sizes only -------------------
def laplace_density(s, beta_param, bias, beta_min=1e-4):
    beta = F32(abs(beta_param) + beta_min)
    x = (s - F32(bias)).astype(F32)
    return (F32(1) / beta) * (F32(0.5) + F32(0.5) * np.sign(x) * np.expm1(-np.abs(x) / beta))


def trilinear_zeros(vol, i0, f, sizes):
    """vol (C, Z, Y, X); i0/f triples (x,y,z) of arrays shape P -> (C, *P), zeros padding."""
    X, Y, Z = sizes
    out = np.zeros((vol.shape[0],) + i0[0].shape, F32)
    for dz in (0, 1):
        for dy in (0, 1):
            for dx in (0, 1):
                xx, yy, zz = i0[0] + dx, i0[1] + dy, i0[2] + dz
                inb = (xx >= 0) & (xx < X) & (yy >= 0) & (yy < Y) & (zz >= 0) & (zz < Z)
                wgt = (f[0] if dx else F32(1) - f[0]) * (f[1] if dy else F32(1) - f[1]) * \
                      (f[2] if dz else F32(1) - f[2])
                wgt = np.where(inb, wgt, F32(0)).astype(F32)
                out += vol[:, np.clip(zz, 0, Z - 1), np.clip(yy, 0, Y - 1), np.clip(xx, 0, X - 1)] * wgt
    return out


def render_camera_sequential(geom, vol_cam, lo, ext, mids, bg_depth, beta_param, bias):
    """One sample: geom (N,D,h,w,3) nan_to_num'ed; vol_cam (1+K+3, Z,Y,X).
    Front-to-back per-ray recurrences.  -> rgb (N,3,h,w), seg (N,K,h,w), depth (N,1,h,w)."""
    N, D, h, w, _ = geom.shape
    nch, Z, Y, X = vol_cam.shape
    K = nch - 4
    tau = np.zeros((N, h, w), F32)
    acc = np.zeros((N, h, w), F32)
    dep = np.zeros((N, h, w), F32)
    ch = np.zeros((nch - 1, N, h, w), F32)
    for i in range(D - 1):
        idx = render_indices(geom[:, i], lo, ext, (X, Y, Z))
        v = trilinear_zeros(vol_cam, idx["i0"], idx["f"], (X, Y, Z)) * idx["mask"][None]
        v = nan_to_num(v)
        sigma = laplace_density(v[0], beta_param, bias)
        dlt = geom[:, i + 1] - geom[:, i]
        delta = np.sqrt((dlt * dlt).sum(-1)).astype(F32)
        sd = sigma * delta
        wgt = (F32(1) - np.exp(-sd)) * np.exp(-tau)
        acc += wgt
        dep += wgt * F32(mids[i])
        ch += wgt[None] * v[1:]
        tau = tau + sd
    depth = dep + (F32(1) - acc) * F32(bg_depth)
    seg = np.moveaxis(ch[:K], 0, 1)
    rgb = np.moveaxis(ch[K:K + 3], 0, 1)
    return rgb, seg, depth[:, None]
