"""The C-ABI library loads and exports every symbol include/lane_tracker_b200.h declares; struct layouts
of the ctypes binding match the C header (no compute calls: this runs without a GPU)."""
import ctypes as C
import os
import re
import subprocess
import sys
import zlib

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "lane_tracker_b200.h")


@pytest.fixture(scope="module")
def lib():
    from lane_tracker_b200 import build, _lib
    if not os.path.exists(_lib.LIB_PATH):
        build.build()
    return _lib.load()


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(lt_[a-z_0-9]+)\s*\(", src)))


def test_exports_every_declared_symbol(lib):
    from lane_tracker_b200 import _lib
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), "library does not export %s" % n
        assert n in _lib.SIGNATURES, "binding does not cover %s" % n
    assert lib.lt_abi_version() == _lib.LT_ABI_VERSION == 4


def test_struct_layouts_match_header(tmp_path, lib):
    from lane_tracker_b200 import _lib
    prog = tmp_path / "sz.c"
    prog.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "lane_tracker_b200.h"\n'
                    'int main(){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(lt_config), sizeof(lt_params),'
                    'sizeof(lt_result), sizeof(lt_state), offsetof(lt_result, left_fit), offsetof(lt_state, last_left),'
                    'offsetof(lt_params, partial), sizeof(lt_vis), offsetof(lt_vis, left_fit));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(prog), "-o", str(exe)])
    got = [int(v) for v in subprocess.check_output([str(exe)]).split()]
    want = [C.sizeof(_lib.lt_config), C.sizeof(_lib.lt_params), C.sizeof(_lib.lt_result), C.sizeof(_lib.lt_state),
            _lib.lt_result.left_fit.offset, _lib.lt_state.last_left.offset, _lib.lt_params.partial.offset,
            C.sizeof(_lib.lt_vis), _lib.lt_vis.left_fit.offset]
    assert got == want


def test_default_params_are_the_reference_defaults(lib):
    from lane_tracker_b200 import _lib
    from lane_tracker_b200.tracker import make_params, PROCESS_DEFAULTS
    p = _lib.lt_params()
    lib.lt_default_params(C.byref(p))
    q = make_params()
    for name, _ in _lib.lt_params._fields_:
        assert getattr(p, name) == getattr(q, name), name
    assert (p.ksize_r, p.C_r, p.ksize_b, p.C_b, p.bandwidth, p.partial, p.n_tries) == (15, 8, 35, 5, 25, 1.0, 2)
    assert PROCESS_DEFAULTS["no_success_limit"] == 8 and PROCESS_DEFAULTS["ignore_sides"] == 360
    with pytest.raises(ValueError, match="Unexpected filter mode"):
        make_params(filter_type="gaussian")


def test_literal_lab_tables_match_the_oracle():
    from oracle.cvops import lab_tables
    g, cb = lab_tables()
    src = open(os.path.join(ROOT, "lane_tracker_b200", "csrc", "lab_tables.inc")).read()
    nums = [int(v) for v in re.findall(r"\b\d+\b", re.sub(r"//.*", "", src).replace("[256]", "").replace("[3072]", ""))]
    assert nums[:256] == [int(v) for v in g] and nums[256:] == [int(v) for v in cb]
    assert zlib.crc32(g.astype("<u2").tobytes()) == 0xb9e3b8e8 and zlib.crc32(cb.astype("<u2").tobytes()) == 0xaa0d61cf


def test_product_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from lane_tracker_b200 import LaneTracker, _lib, synth
    with pytest.raises(_lib.LaneTrackerError):
        LaneTracker(**synth.shipped_calibration())
    import numpy as np
    from lane_tracker_b200 import create_split_view            # the split-view helper resizes on the GPU: no CPU path
    with pytest.raises(_lib.LaneTrackerError):
        create_split_view((64, 64), [np.zeros((32, 32, 3), np.uint8)], [(0, 0)], [(16, 16)])


def test_product_never_imports_the_oracle():
    code = ("import sys; import lane_tracker_b200, lane_tracker_b200.tracker, lane_tracker_b200.synth; "
            "print(any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules))")
    out = subprocess.check_output([sys.executable, "-c", code], cwd=ROOT).decode().strip()
    assert out == "False"
    for fn in os.listdir(os.path.join(ROOT, "lane_tracker_b200")):
        if fn.endswith(".py"):
            assert "oracle" not in open(os.path.join(ROOT, "lane_tracker_b200", fn)).read().replace("# oracle", "")
