"""ORACLE (test infrastructure only): import the real reference backbone with three stubs.

Only usable where /root/reference exists (the build container).  The GPU box never has it;
tests that need it are skipped there and rely on ``tests/golden`` instead.  Recipe: SURVEY B.1.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types
import warnings

REFERENCE_ROOT = "/root/reference"
_BV2 = os.path.join(REFERENCE_ROOT, "src/layers/backbones/base_vampire2.py")


def reference_available() -> bool:
    return os.path.isfile(_BV2)


def load_reference_module():
    import torch.nn as nn

    class _Dummy(nn.Module):
        def init_weights(self):
            pass

    for n in ["mmdet3d", "mmdet3d.models", "mmdet", "mmdet.models", "matplotlib", "matplotlib.pyplot"]:
        if n not in sys.modules:
            sys.modules[n] = types.ModuleType(n)
    sys.modules["mmdet3d.models"].build_neck = lambda *a, **k: _Dummy()
    sys.modules["mmdet.models"].build_backbone = lambda *a, **k: _Dummy()
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    spec = importlib.util.spec_from_file_location("ref_bv2", _BV2)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def reference_bilinear_get_voxel_feats():
    """The unbound ``BaseBiLinear.get_voxel_feats`` (base_bilinear.py:471): it only touches ``self.get_pixel``,
    ``self.final_dim`` and ``self.vZ/vY/vX``, all of which a ``BaseVAMPIRE2`` instance provides identically."""
    load_reference_module()          # installs the import stubs
    path = os.path.join(REFERENCE_ROOT, "src/layers/backbones/base_bilinear.py")
    spec = importlib.util.spec_from_file_location("ref_bilinear", path)
    mod = importlib.util.module_from_spec(spec)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        spec.loader.exec_module(mod)
    return mod.BaseBiLinear.get_voxel_feats


def build_reference_backbone(conf: dict):
    """Instantiate ``BaseVAMPIRE2`` with dummy image encoder (never executed)."""
    mod = load_reference_module()
    kw = dict(conf)
    kw.setdefault("img_backbone_conf", {})
    kw.setdefault("img_neck_conf", dict(out_channels=[8] * 4))
    kw.setdefault("output_channels", 80)
    kw.setdefault("cat_pos", True)
    kw.pop("num_cams", None)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        bb = mod.BaseVAMPIRE2(**kw)
    return bb
