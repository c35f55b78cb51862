"""Build liblane_tracker_b200.so in-tree with nvcc for sm_100a (no torch extension machinery needed:
the library exposes a plain C ABI, include/lane_tracker_b200.h)."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
BUILD = os.path.join(HERE, "_build")
LIB = os.path.join(HERE, "liblane_tracker_b200.so")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xptxas", "-v"]
# fp64 expressions whose rounding must match OpenCV / NumPy live in these files: no FMA contraction
SOURCES = {
    "lt_api.cu": [],
    "lt_remap.cu": ["-fmad=false"],
    "lt_filter.cu": [],
    "lt_morph.cu": [],
    "lt_search.cu": ["-fmad=false"],
    "lt_vis.cu": ["-fmad=false"],
}


def nvcc_path():
    p = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(p):
        raise RuntimeError("nvcc not found: the lane_tracker_b200 CUDA library cannot be built")
    return p


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    os.makedirs(BUILD, exist_ok=True)
    nvcc = nvcc_path()
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".inc", ".h"))]
    headers.append(os.path.join(HERE, "..", "include", "lane_tracker_b200.h"))
    headers.append(os.path.abspath(__file__))
    objs = []
    logs = []
    for src, extra in SOURCES.items():
        s = os.path.join(CSRC, src)
        o = os.path.join(BUILD, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _stale(o, [s] + headers):
            cmd = [nvcc] + ARCH + COMMON + extra + ["-c", s, "-o", o]
            r = subprocess.run(cmd, capture_output=True, text=True)
            logs.append(r.stderr)
            if r.returncode != 0:
                sys.stderr.write(r.stdout + r.stderr)
                raise RuntimeError("nvcc failed on %s" % src)
            with open(o + ".ptxas.log", "w") as f:
                f.write(r.stderr)
            if verbose:
                sys.stderr.write(r.stderr)
    if force or _stale(LIB, objs):
        cmd = [nvcc] + ARCH + ["-shared", "-o", LIB] + objs
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
