"""CPU: the 'affine ida' shortcuts of the fused kernels are VALUE-IDENTICAL to the reference's full 4x4 chains.

vb_common.cuh collapses the first mat-vec of get_geometry (frustum_point_affine) and the last one of get_pixel
(project_voxel_affine) when ida is a 2-D affine map and the homogeneous rows are e_w.  Here the collapsed forms
are written out in NumPy fp32 (one rounding per operation, no FMA -- the same order the kernels use with
__fmul_rn/__fadd_rn) and compared for equality of VALUES with oracle/strict_np.py's full chains, which
tests/test_oracle_golden.py pins bit for bit to the reference.  (+0 / -0 may differ; no compare or floor
downstream can tell them apart.)"""
import numpy as np
import pytest

from oracle import strict_np as sn
from vampire_b200 import synth
from vampire_b200.config import MINI
from vampire_b200.lattice import build_lattice
from vampire_b200.matrices import prepare_matrices

F32 = np.float32


def _prep(mode, seed):
    m = synth.make_mats(MINI, 2, mode, seed=seed)
    return prepare_matrices(m["sensor2ego_mats"][:, 0], m["intrin_mats"][:, 0], m["ida_mats"][:, 0], m["bda_mat"]).numpy()


def _is_affine(prep):
    ida, ida_inv, ke, bi = prep[:, :, 2], prep[:, :, 3], prep[:, :, 1], prep[:, :, 0]
    e_z, e_w = np.array([0, 0, 1, 0], F32), np.array([0, 0, 0, 1], F32)
    ok = True
    for M in (ida, ida_inv):
        ok &= bool((M[..., 0, 2] == 0).all() and (M[..., 1, 2] == 0).all() and (M[..., 2, :] == e_z).all() and (M[..., 3, :] == e_w).all())
    return ok and bool((ke[..., 3, :] == e_w).all() and (bi[..., 3, :] == e_w).all())


@pytest.mark.parametrize("mode,seed", [("val", 1), ("train", 2), ("stress", 3), ("stress", 4)])
def test_affine_forms_equal_the_full_chains(mode, seed):
    prep = _prep(mode, seed)
    assert _is_affine(prep), "the dataset's ida construction (nusc_det_seg_dataset.py:118-146) is a 2-D affine map"
    lat = build_lattice(MINI)
    us, vs, ds = lat.us.numpy(), lat.vs.numpy(), lat.ds.numpy()
    xs, ys, zs = lat.xs.numpy(), lat.ys.numpy(), lat.zs.numpy()

    # ---- G2: frustum_point_affine ----
    full = sn.frustum_points(prep, us, vs, ds)                                     # (B,N,D,fH,fW,3)
    I = prep[:, :, 3][:, :, None, None, None]                                      # ida^-1
    E = prep[:, :, 4][:, :, None, None, None]                                      # E.K^-1
    Bd = prep[:, :, 5][:, :, None, None, None]                                     # bda
    U = us.astype(F32)[None, None, None, None, :]
    V = vs.astype(F32)[None, None, None, :, None]
    Dd = ds.astype(F32)[None, None, :, None, None]
    A0 = (I[..., 0, 0] * U + I[..., 0, 1] * V) + I[..., 0, 3]
    A1 = (I[..., 1, 0] * U + I[..., 1, 1] * V) + I[..., 1, 3]
    p0, p1 = A0 * Dd, A1 * Dd
    q = [((E[..., i, 0] * p0 + E[..., i, 1] * p1) + E[..., i, 2] * Dd) + E[..., i, 3] for i in range(4)]
    r = sn._mv(Bd, q)
    fast = np.stack(np.broadcast_arrays(*r[:3]), axis=-1).astype(F32)
    assert np.array_equal(fast, full)            # array_equal: +0 == -0, any other difference fails

    # ---- G1: project_voxel_affine ----
    full = sn.project_voxels(prep, xs, ys, zs)                                     # (B,N,Z,Y,X,3)
    X = xs.astype(F32)[None, None, None, None, :]
    Y = ys.astype(F32)[None, None, None, :, None]
    Z = zs.astype(F32)[None, None, :, None, None]
    Bi = prep[:, :, 0][:, :, None, None, None]
    KE = prep[:, :, 1][:, :, None, None, None]
    Id = prep[:, :, 2][:, :, None, None, None]
    p = [((Bi[..., i, 0] * X + Bi[..., i, 1] * Y) + Bi[..., i, 2] * Z) + Bi[..., i, 3] for i in range(3)]
    q = [((KE[..., i, 0] * p[0] + KE[..., i, 1] * p[1]) + KE[..., i, 2] * p[2]) + KE[..., i, 3] for i in range(3)]
    zc = np.maximum(q[2], F32(1e-6))
    u, v = q[0] / zc, q[1] / zc
    fast = np.stack(np.broadcast_arrays((Id[..., 0, 0] * u + Id[..., 0, 1] * v) + Id[..., 0, 3],
                                        (Id[..., 1, 0] * u + Id[..., 1, 1] * v) + Id[..., 1, 3], q[2]), axis=-1).astype(F32)
    assert np.array_equal(fast, full)
