#!/usr/bin/env python3
"""Build tuning variants of the CUDA library: tools/build_variants.py name=-DFLAG[,-DFLAG...] ...
-> lane_tracker_b200/_variants/liblane_tracker_b200_<name>.so (selected with LT_LIBRARY_VARIANT=<name>)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lane_tracker_b200 import build as B  # noqa: E402


def main():
    out_dir = os.path.join(B.HERE, "_variants")
    os.makedirs(out_dir, exist_ok=True)
    for spec in sys.argv[1:]:
        name, flags = spec.split("=", 1)
        flags = [f for f in flags.split(",") if f]
        bdir = os.path.join(out_dir, "_build_" + name)
        os.makedirs(bdir, exist_ok=True)
        objs = []
        procs = []
        for src, extra in B.SOURCES.items():
            o = os.path.join(bdir, src.replace(".cu", ".o"))
            objs.append(o)
            cmd = [B.nvcc_path()] + B.ARCH + [c for c in B.COMMON if c not in ("-Xptxas", "-v")] + extra + flags + ["-c", os.path.join(B.CSRC, src), "-o", o]
            procs.append(subprocess.Popen(cmd))
        for p in procs:
            if p.wait() != 0:
                raise SystemExit("nvcc failed for variant " + name)
        lib = os.path.join(out_dir, "liblane_tracker_b200_%s.so" % name)
        subprocess.check_call([B.nvcc_path()] + B.ARCH + ["-shared", "-o", lib] + objs)
        print(lib)


if __name__ == "__main__":
    main()
