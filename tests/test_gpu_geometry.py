"""GPU: G1/G2/L2/R2 -- geometry, masks and corner indices, BIT-EXACT against the reference
(golden SHA-256 digests) and against the strict NumPy oracle, through the C ABI."""
import numpy as np
import pytest
import torch

from helpers import CASES, Case, sha
from oracle import strict_np as sn

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", params=list(CASES))
def case(request):
    return Case(request.param)


def _ops(case):
    from vampire_b200 import ops
    return ops, ops.register_config(case.cfg)


def test_get_pixel_bits(case):
    ops, cid = _ops(case)
    pix = ops.get_pixel(case.prep.cuda(), cid, True).cpu().numpy()
    assert sha(pix) == str(case.gold["pix_sha"]), "get_pixel differs from the reference"
    lat = case.lat
    ref = sn.project_voxels(case.gold["prep"], lat.xs.numpy(), lat.ys.numpy(), lat.zs.numpy())
    assert np.array_equal(pix.view(np.uint32), ref.view(np.uint32))


def test_get_geometry_bits(case):
    ops, cid = _ops(case)
    geom = ops.get_geometry(case.prep.cuda(), cid, True, False).cpu().numpy()
    assert sha(geom) == str(case.gold["geom_sha"]), "get_geometry differs from the reference"


def test_lift_indices_bits(case):
    ops, cid = _ops(case)
    valid, i0, frac = ops.lift_indices(case.prep.cuda(), cid, True)
    valid, i0, frac = valid.cpu().numpy(), i0.cpu().numpy(), frac.cpu().numpy()
    assert int(valid.sum()) == int(case.gold["lift_valid_count"])
    assert sha(valid) == str(case.gold["lift_valid_sha"])
    assert sha(i0) == str(case.gold["lift_i0_sha"])
    assert sha(i0[valid.astype(bool)]) == str(case.gold["lift_i0_valid_sha"])
    lat, cfg = case.lat, case.cfg
    pix = sn.project_voxels(case.gold["prep"], lat.xs.numpy(), lat.ys.numpy(), lat.zs.numpy())
    li = sn.lift_indices(pix, cfg.final_dim, cfg.d_bound, (cfg.fW, cfg.fH, cfg.D))
    assert np.array_equal(frac.view(np.uint32), np.stack(li["f"], -1).view(np.uint32))


@pytest.mark.parametrize("from_tensor", [False, True])
def test_render_indices_bits(case, from_tensor):
    ops, cid = _ops(case)
    prep = case.prep.cuda()
    geom = ops.get_geometry(prep, cid, True, True) if from_tensor else None
    mask, i0, frac = ops.render_indices(prep, cid, True, geom)
    mask, i0 = mask.cpu().numpy(), i0.cpu().numpy()
    assert int(mask.sum()) == int(case.gold["render_mask_count"])
    assert sha(mask) == str(case.gold["render_mask_sha"])
    assert sha(i0[mask.astype(bool)]) == str(case.gold["render_i0_masked_sha"])


def test_no_bda_path_matches_oracle():
    """mats_dict without 'bda_mat' skips the bda products entirely (BV2:343, 370)."""
    from vampire_b200 import ops, synth
    from vampire_b200.config import MINI
    from vampire_b200.lattice import build_lattice
    from vampire_b200.matrices import prepare_matrices
    cid = ops.register_config(MINI)
    lat = build_lattice(MINI)
    m = synth.make_mats(MINI, 2, "stress", seed=9)
    prep = prepare_matrices(m["sensor2ego_mats"][:, 0], m["intrin_mats"][:, 0], m["ida_mats"][:, 0], None)
    pix = ops.get_pixel(prep.cuda(), cid, False).cpu().numpy()
    ref = sn.project_voxels(prep.numpy(), lat.xs.numpy(), lat.ys.numpy(), lat.zs.numpy(), has_bda=False)
    assert np.array_equal(pix.view(np.uint32), ref.view(np.uint32))
    geom = ops.get_geometry(prep.cuda(), cid, False, False).cpu().numpy()
    refg = sn.frustum_points(prep.numpy(), lat.us.numpy(), lat.vs.numpy(), lat.ds.numpy(), has_bda=False)
    assert np.array_equal(geom.view(np.uint32), refg.view(np.uint32))


def test_degenerate_geometry_nan_to_num():
    """A singular ida makes get_geometry emit inf/NaN; nan_to_num must map them like torch (BV2:612)."""
    from vampire_b200 import ops
    from vampire_b200.config import MINI
    cid = ops.register_config(MINI)
    prep = torch.eye(4).expand(1, 6, 6, 4, 4).clone()
    prep[0, 0, 3] = float("nan")
    prep[0, 1, 4, 0, 0] = float("inf")
    g = ops.get_geometry(prep.cuda(), cid, True, True).cpu()
    assert torch.isfinite(g).all()
    assert (g[0, 0] == -1e3).all()
    raw = ops.get_geometry(prep.cuda(), cid, True, False).cpu()
    assert torch.equal(torch.nan_to_num(raw, -1e3), g)


@pytest.mark.parametrize("seed", list(range(100, 112)))
@pytest.mark.parametrize("mode", ["train", "stress"])
def test_random_rigs_bit_exact(seed, mode):
    """Property: for ANY camera rig (random resize/crop, ida rotation + flip, bda rotation/scale/flip)
    coordinates, masks and corner indices equal the strict oracle bit for bit."""
    from vampire_b200 import ops, synth
    from vampire_b200.config import MINI
    from vampire_b200.lattice import build_lattice
    from vampire_b200.matrices import prepare_matrices
    cfg = MINI
    cid = ops.register_config(cfg)
    lat = build_lattice(cfg)
    m = synth.make_mats(cfg, 2, mode, seed=seed)
    prep = prepare_matrices(m["sensor2ego_mats"][:, 0], m["intrin_mats"][:, 0], m["ida_mats"][:, 0], m["bda_mat"])
    pn = prep.numpy()
    dev = prep.cuda()
    pix = sn.project_voxels(pn, lat.xs.numpy(), lat.ys.numpy(), lat.zs.numpy())
    assert np.array_equal(ops.get_pixel(dev, cid, True).cpu().numpy().view(np.uint32), pix.view(np.uint32))
    geom = sn.frustum_points(pn, lat.us.numpy(), lat.vs.numpy(), lat.ds.numpy())
    assert np.array_equal(ops.get_geometry(dev, cid, True, False).cpu().numpy().view(np.uint32), geom.view(np.uint32))
    li = sn.lift_indices(pix, cfg.final_dim, cfg.d_bound, (cfg.fW, cfg.fH, cfg.D))
    valid, i0, frac = ops.lift_indices(dev, cid, True)
    assert np.array_equal(valid.cpu().numpy().astype(bool), li["valid"])
    assert np.array_equal(i0.cpu().numpy(), np.stack(li["i0"], -1).astype(np.int16))
    lo = (cfg.x_bound_seg[0], cfg.y_bound_seg[0], cfg.z_bound_seg[0])
    ext = (cfg.x_bound_seg[1] - lo[0], cfg.y_bound_seg[1] - lo[1], cfg.z_bound_seg[1] - lo[2])
    ri = sn.render_indices(sn.nan_to_num(geom, -1e3)[:, :, :-1], lo, ext, (cfg.vX, cfg.vY, cfg.vZ))
    mask, r0, _ = ops.render_indices(dev, cid, True, None)
    mask = mask.cpu().numpy().astype(bool)
    assert np.array_equal(mask, ri["mask"])
    assert np.array_equal(r0.cpu().numpy()[mask], np.stack(ri["i0"], -1)[ri["mask"]].astype(np.int16))
