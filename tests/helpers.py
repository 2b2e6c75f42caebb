"""Shared test plumbing: golden fixtures, seeded cases, tolerances."""
from __future__ import annotations

import hashlib
import os

import numpy as np
import torch

from vampire_b200 import synth
from vampire_b200.config import MINI, R50_256x704, PathConfig
from vampire_b200.lattice import build_lattice

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

CASES = {
    # name: (config, batch, mats mode, density field)   -- must match oracle/gen_golden.py:main
    "mini_val": (MINI, 2, "val", "random"),
    "mini_stress": (MINI, 2, "stress", "surface"),
    "r50_val_digest": (R50_256x704, 1, "val", "surface"),
    "r50_stress_digest": (R50_256x704, 1, "stress", "random"),
}


def sha(a) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def load_golden(name):
    return np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)


class Case:
    """Seeded inputs of one golden case, regenerated (not stored) and checksum-guarded."""

    def __init__(self, name: str):
        self.name = name
        self.cfg, self.batch, self.mode, self.field = CASES[name]
        self.gold = load_golden(name)
        self.conf = self.cfg.backbone_kwargs()
        self.lat = build_lattice(self.cfg)
        self.mats = synth.make_mats(self.cfg, self.batch, self.mode)
        self.depth, self.ctx = synth.make_lift_inputs(self.cfg, self.batch)
        self.den, self.sem, self.feat, self.rgb = synth.make_render_inputs(self.cfg, self.batch, field=self.field)
        chk = np.array([t.double().sum().item() for t in (self.depth, self.ctx, self.den, self.sem, self.feat, self.rgb)])
        self.inputs_match_golden = bool(np.allclose(chk, self.gold["in_checksum"], rtol=1e-7, atol=0))
        self.prep = torch.from_numpy(self.gold["prep"])  # the build container's prepared matrices
        self.stride = int(self.gold["meta_stride"])

    def mat_args(self):
        m = self.mats
        return m["sensor2ego_mats"][:, 0], m["intrin_mats"][:, 0], m["ida_mats"][:, 0], m["bda_mat"]

    def seg_lo_ext(self):
        c = self.cfg
        lo = (c.x_bound_seg[0], c.y_bound_seg[0], c.z_bound_seg[0])
        ext = (c.x_bound_seg[1] - c.x_bound_seg[0], c.y_bound_seg[1] - c.y_bound_seg[0],
               c.z_bound_seg[1] - c.z_bound_seg[0])
        return lo, ext

    def cotangents(self):
        c, B = self.cfg, self.batch
        shapes = [(B, c.C, c.vZ, c.vY, c.vX), (B, c.num_cams, 3, c.fH, c.fW), (B, c.num_cams, c.K, c.fH, c.fW),
                  (B, c.num_cams, 1, c.fH, c.fW), (B, 3, c.oY, c.oX), (B, c.K, c.oY, c.oX), (B, 1, c.oY, c.oX),
                  (B, 1, c.oZ, c.oY, c.oX), (B, c.C, c.oZ, c.oY, c.oX)]
        return synth.make_cotangents(shapes)


def golden_value(gold, key, full_array):
    """Return (expected, got) aligned: fixtures hold either the full array or a strided sample."""
    if key in gold.files:
        return gold[key], np.asarray(full_array)
    stride = int(gold["meta_stride"])
    return gold[key + "_strided"], np.asarray(full_array).reshape(-1)[::stride]


ACHIEVED = {}      # what -> worst (max abs err / scale) seen in this session; printed at the end (conftest.py)


def assert_close_scaled(got, exp, rel, what="", scale=None):
    """|got - exp| <= rel * max|exp|  (the inf-norm-relative bar of SURVEY B.3) plus elementwise rtol.
    `scale` overrides max|exp| when `exp` is only a strided sample of the full tensor."""
    got = np.asarray(got, dtype=np.float64)
    exp = np.asarray(exp, dtype=np.float64)
    assert got.shape == exp.shape, f"{what}: shape {got.shape} vs {exp.shape}"
    scale = max(np.abs(exp).max(), 1e-30) if scale is None else float(scale)
    err = np.abs(got - exp)
    if err.size:
        a = float(np.nanmax(err)) / scale
        prev = ACHIEVED.get((what, rel))
        ACHIEVED[(what, rel)] = max(a, prev) if prev is not None else a
    bad = err > rel * scale + rel * np.abs(exp)
    assert not bad.any(), (f"{what}: {int(bad.sum())} / {bad.size} elements off; max abs err {err.max():.3e} "
                           f"(scale {scale:.3e}, allowed {rel * scale:.3e})")
