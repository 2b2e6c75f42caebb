import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def pytest_terminal_summary(terminalreporter):
    """Achieved max errors (relative to max|reference|) per compared quantity, next to the tolerance asserted."""
    try:
        from helpers import ACHIEVED
    except Exception:
        return
    if not ACHIEVED:
        return
    terminalreporter.section("achieved max |err| / max|ref|  (tolerance asserted)")
    for (what, rel), a in sorted(ACHIEVED.items()):
        terminalreporter.write_line(f"  {what:48s} {a:9.2e}   ({rel:.0e})")
