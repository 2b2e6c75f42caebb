"""CPU: exhaustive proof-by-enumeration that the constant-divisor division of the bit-exact zone
(vampire_b200/csrc/vb_common.cuh ``sdiv_const``: RN(1/y) from the host + two FMA corrections) returns the
IEEE-754 quotient for EVERY fp32 mantissa, both signs, exponents 2^-100 .. 2^100 -- for each divisor the
shipped configurations use (W-1, H-1, d_bound extent, seg-grid extents).  tests/native/div_const_check.c."""
import os
import shutil
import subprocess

import numpy as np
import pytest

from vampire_b200 import cabi
from vampire_b200.config import MINI, R50_256x704, R50_512x1408

SRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "native", "div_const_check.c")


def _divisors():
    out = set()
    for cfg in (R50_256x704, R50_512x1408, MINI):
        g = cabi.make_grid(cfg, 1)
        out.update(float(np.float32(v)) for v in (g.img_w_m1, g.img_h_m1, g.d_ext, g.seg_ext[0], g.seg_ext[1], g.seg_ext[2]))
    return sorted(out)


@pytest.mark.skipif(shutil.which("gcc") is None, reason="needs gcc")
def test_constant_divisor_division_is_correctly_rounded(tmp_path):
    exe = str(tmp_path / "divcheck")
    base = ["gcc", "-O2", "-ffp-contract=off", SRC, "-lm", "-o", exe]
    if subprocess.run(base[:3] + ["-mfma"] + base[3:], capture_output=True).returncode != 0:
        subprocess.run(base, check=True)                       # no hardware FMA: glibc's exact fmaf (slower)
    divs = _divisors()
    assert len(divs) >= 6
    res = subprocess.run([exe] + ["%.9g" % d for d in divs], capture_output=True, text=True, timeout=900)
    lines = [l.split() for l in res.stdout.strip().splitlines()]
    assert len(lines) == len(divs), res.stdout + res.stderr
    for (y, bad), d in zip(lines, divs):
        assert float(y) == pytest.approx(d, rel=1e-7) and int(bad) == 0, f"divisor {y}: {bad} quotients differ from IEEE"
    assert res.returncode == 0
