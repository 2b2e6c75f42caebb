"""CPU: the C-ABI library loads and exports every symbol include/vb200.h declares (no compute
calls without a GPU), plus the host-side logic (config, lattice, matrices, synth, grid constants)."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch


from oracle import torch_path as tp
from vampire_b200 import cabi, synth
from vampire_b200.config import MINI, R50_256x704, R50_512x1408, PathConfig
from vampire_b200.lattice import build_lattice
from vampire_b200.matrices import prepare_matrices

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(REPO, "include", "vb200.h")).read()
    declared = sorted(set(re.findall(r"\b(vb200_[a-z_0-9]+)\s*\(", header)))
    assert declared, "no declarations parsed"
    lib = cabi.lib()
    for name in declared:
        assert hasattr(lib, name), f"libvb200.so does not export {name}"
    assert sorted(cabi.exported_symbols()) == declared
    assert lib.vb200_version() == 203
    assert b"sm_100" in lib.vb200_strerror(-3)


def test_no_gpu_means_loud_failure():
    if torch.cuda.is_available():
        pytest.skip("has a GPU")
    from vampire_b200 import ops
    cid = ops.register_config(MINI)
    mats = torch.eye(4).expand(1, 6, 6, 4, 4).contiguous()
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.get_pixel(mats, cid, True)


def test_struct_layout_matches_header():
    """ctypes mirrors of the header structs: field order and count."""
    header = open(os.path.join(REPO, "include", "vb200.h")).read()
    body = re.search(r"typedef struct VbGrid \{(.*?)\} VbGrid;", header, re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        decl = re.sub(r"^(int32_t|int16_t|float)\s+", "", decl)
        names += [re.sub(r"\[\d+\]", "", n.strip()) for n in decl.split(",")]
    assert names == [f[0] for f in cabi.VbGrid._fields_]
    assert cabi.VbGrid.seg_lo.size == 12 and cabi.VbGrid.seg_ext.size == 12
    # a by-value kernel parameter: has_bda / density_mode share one word so that it stays at 128 bytes (vb200.h)
    assert C.sizeof(cabi.VbGrid) == 128 and cabi.VbGrid.density_mode.offset == cabi.VbGrid.has_bda.offset + 2


def test_config_sizes():
    c = R50_256x704
    assert (c.fH, c.fW, c.D, c.S) == (64, 176, 86, 85)
    assert (c.vZ, c.vY, c.vX, c.oZ, c.oY, c.oX) == (20, 256, 256, 10, 256, 256)
    assert (c.cam_channels, c.all_channels) == (22, 38)
    assert (R50_512x1408.fH, R50_512x1408.fW) == (128, 352)
    assert PathConfig.from_backbone_conf(c.backbone_kwargs()) == c


@pytest.mark.parametrize("cfg", [MINI, R50_256x704])
def test_lattice_equals_reference_buffers(cfg):
    lat = build_lattice(cfg)
    buf = tp.build_buffers(cfg.backbone_kwargs())
    assert torch.equal(buf["frustum"][0, 0, :, 0], lat.us)
    assert torch.equal(buf["frustum"][0, :, 0, 1], lat.vs)
    assert torch.equal(buf["frustum"][:, 0, 0, 2], lat.ds)
    assert torch.equal(buf["voxel_coords"][0, 0, :, 0], lat.xs)
    assert torch.equal(buf["voxel_coords"][0, :, 0, 1], lat.ys)
    assert torch.equal(buf["voxel_coords"][:, 0, 0, 2], lat.zs)
    assert torch.equal(buf["output_coords"][0, 0, :, 0], lat.oxs)
    assert torch.equal(buf["output_coords"][:, 0, 0, 2], lat.ozs)
    assert torch.equal(buf["camera_mids"], lat.mids)
    assert torch.equal(buf["bev_mids"], lat.bev_mids)
    assert lat.packed().numel() == sum(v.numel() for v in lat.__dict__.values())


def test_grid_constants_round_like_torch():
    g = cabi.make_grid(R50_256x704, 3, True)
    assert (g.B, g.N, g.D, g.C, g.K) == (3, 6, 86, 16, 18)
    assert np.float32(g.d_ext) == np.float32(70.4 - 2.0)
    assert np.float32(g.d_hi) == np.float32(70.4)
    assert np.float32(g.seg_ext[0]) == np.float32(51.2 - (-51.2))
    assert g.x_hi == 703.5 and g.img_w_m1 == 703.0 and g.y_hi == 255.5


def test_prepare_matrices_slots():
    mats = synth.make_mats(MINI, 2, "stress")
    a = (mats["sensor2ego_mats"][:, 0], mats["intrin_mats"][:, 0], mats["ida_mats"][:, 0], mats["bda_mat"])
    p = prepare_matrices(*a)
    assert p.shape == (2, 6, 6, 4, 4)
    assert torch.equal(p[:, :, 2], a[2])
    assert torch.equal(p[:, :, 1], a[1].matmul(torch.inverse(a[0])))
    assert torch.allclose(p[:, :, 0] @ p[:, :, 5], torch.eye(4).expand(2, 6, 4, 4), atol=1e-5)
    p0 = prepare_matrices(a[0], a[1], a[2], None)
    assert torch.equal(p0[:, :, 0], torch.eye(4).expand(2, 6, 4, 4))


def test_synth_is_deterministic():
    a = synth.make_mats(MINI, 2, "stress")
    b = synth.make_mats(MINI, 2, "stress")
    for k in a:
        assert torch.equal(a[k], b[k])
    d1, c1 = synth.make_lift_inputs(MINI, 1)
    d2, c2 = synth.make_lift_inputs(MINI, 1)
    assert torch.equal(d1, d2) and torch.equal(c1, c2)
    assert torch.allclose(d1.sum(2), torch.ones(1, 6, MINI.fH, MINI.fW), atol=1e-5)
