"""Build libvb200.so in-tree with nvcc for sm_100a (replaces the reference's empty
``build_ext`` in /root/reference/setup.py:10-26).

    python -m vampire_b200.build [--force] [--verbose]

The library lands in ``vampire_b200/_lib/libvb200.so`` (git-ignored, but shipped to the GPU box
with the repo snapshot).  Objects are cached per source hash so rebuilding after editing one
kernel file recompiles only that file.
"""
from __future__ import annotations

import argparse
import concurrent.futures as cf
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "_lib")
OBJDIR = os.path.join(LIBDIR, "obj")
LIBPATH = os.path.join(LIBDIR, "libvb200.so")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "-Xptxas", "-v",
]


def _nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: libvb200 cannot be built (there is no fallback path)")
    return exe


def _sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest(path: str) -> str:
    h = hashlib.sha256()
    h.update(" ".join(NVCC_FLAGS).encode())
    for dep in [path] + sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")) + \
            [os.path.join(INCLUDE, "vb200.h")]:
        with open(dep, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()[:16]


def _compile(src: str, verbose: bool) -> str:
    obj = os.path.join(OBJDIR, os.path.basename(src)[:-3] + "." + _digest(src) + ".o")
    if os.path.exists(obj):
        return obj
    cmd = [_nvcc()] + NVCC_FLAGS + ["-I", INCLUDE, "-c", src, "-o", obj]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{res.stdout}\n{res.stderr}")
    log = os.path.join(OBJDIR, os.path.basename(src)[:-3] + ".ptxas.log")
    with open(log, "w") as fh:
        fh.write(res.stderr)
    if verbose:
        sys.stderr.write(res.stderr)
    return obj


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJDIR, exist_ok=True)
    srcs = _sources()
    if force:
        for f in os.listdir(OBJDIR):
            os.remove(os.path.join(OBJDIR, f))
    with cf.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(lambda s: _compile(s, verbose), srcs))
    stamp = os.path.join(LIBDIR, "libvb200.stamp")
    want = " ".join(os.path.basename(o) for o in objs)
    if not force and os.path.exists(LIBPATH) and os.path.exists(stamp) and open(stamp).read() == want:
        return LIBPATH
    cmd = [_nvcc(), "-shared", "-o", LIBPATH] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"link failed:\n{res.stdout}\n{res.stderr}")
    with open(stamp, "w") as fh:
        fh.write(want)
    # drop stale objects
    keep = {os.path.basename(o) for o in objs}
    for f in os.listdir(OBJDIR):
        if f.endswith(".o") and f not in keep:
            os.remove(os.path.join(OBJDIR, f))
    return LIBPATH


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    a = ap.parse_args()
    print(build(a.force, a.verbose))
