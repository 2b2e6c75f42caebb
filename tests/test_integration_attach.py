"""CPU, build container only: attaching to a REAL reference backbone keeps its interface and its
state dict (no GPU: only the structure is checked here; numerics are covered by the GPU tests)."""
import inspect

import pytest
import torch

from oracle.ref_import import build_reference_backbone, reference_available
from vampire_b200.config import MINI, PathConfig
from vampire_b200.integration import attach, backbone_conf_of

pytestmark = pytest.mark.skipif(not reference_available(), reason="/root/reference not present")


def test_attach_preserves_interface_and_state_dict():
    bb = build_reference_backbone(MINI.backbone_kwargs())
    keys_before = list(bb.state_dict().keys())
    sigs = {n: inspect.signature(getattr(bb, n)) for n in
            ("get_geometry", "get_pixel", "get_voxel_feats", "volume_rendering_from_multiple_views")}
    attach(bb)
    assert list(bb.state_dict().keys()) == keys_before and "density.beta" in keys_before
    for n, sig in sigs.items():
        assert list(inspect.signature(getattr(bb, n)).parameters) == list(sig.parameters), n
    assert bb._vb200_path.density.beta is bb.density.beta          # shared parameter, not a copy
    assert PathConfig.from_backbone_conf(backbone_conf_of(bb)) == MINI
    assert hasattr(bb, "lift_pool") and hasattr(bb, "render")
    from vampire_b200.view_transform import UpsampleB200
    assert isinstance(bb.upsample2d, UpsampleB200) and bb.upsample2d.scale_factor == MINI.upsample_factor


def test_attached_methods_refuse_cpu_tensors():
    if torch.cuda.is_available():
        pytest.skip("has a GPU")
    bb = attach(build_reference_backbone(MINI.backbone_kwargs()))
    from vampire_b200 import synth
    m = synth.make_mats(MINI, 1)
    with pytest.raises(RuntimeError, match="CUDA"):
        bb.get_pixel(m["sensor2ego_mats"][:, 0], m["intrin_mats"][:, 0], m["ida_mats"][:, 0], m["bda_mat"])
