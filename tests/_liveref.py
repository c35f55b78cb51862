"""Loader for the LIVE reference (only where /root/reference exists, i.e. the build container).

Materialises a patched copy of the reference's two modules into a temporary
directory (never into the repo) and imports it.  The patch is exactly the three
NumPy>=1.24 compatibility casts of SURVEY.md Appendix C, which reproduce the
truncation semantics of the NumPy the reference was written for:

  lane_tracker.py:466  img1[:int(img1.shape[0]*(1-partial)),:] = 0
  lane_tracker.py:518  np.linspace(..., int(img_height*partial))
  lane_tracker.py:528  np.int -> int
"""
import importlib.util
import io
import contextlib
import os
import sys
import tempfile

REFERENCE_DIR = "/root/reference"


def available():
    return os.path.isfile(os.path.join(REFERENCE_DIR, "lane_tracker.py"))


_cached = None


def load():
    global _cached
    if _cached is not None:
        return _cached
    src = open(os.path.join(REFERENCE_DIR, "lane_tracker.py")).read()
    a = "img1[:img1.shape[0]*(1-partial),:] = 0"
    b = "np.linspace(img_height*(1-partial), img_height-1, img_height*partial)"
    assert a in src and b in src and "np.int)" in src
    src = src.replace(a, "img1[:int(img1.shape[0]*(1-partial)),:] = 0")
    src = src.replace(b, "np.linspace(img_height*(1-partial), img_height-1, int(img_height*partial))")
    src = src.replace("astype(np.int)", "astype(int)")
    d = tempfile.mkdtemp(prefix="lt_liveref_")
    with open(os.path.join(d, "lane_tracker.py"), "w") as f:
        f.write(src)
    with open(os.path.join(d, "utils.py"), "w") as f:
        f.write(open(os.path.join(REFERENCE_DIR, "utils.py")).read())
    sys.path.insert(0, d)
    try:
        spec = importlib.util.spec_from_file_location("lt_liveref_lane_tracker", os.path.join(d, "lane_tracker.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        sys.path.remove(d)
    _cached = mod
    return mod


def make_tracker(**kw):
    import pickle
    mod = load()
    cal = pickle.load(open(os.path.join(REFERENCE_DIR, "cam_calib.p"), "rb"))
    wp = pickle.load(open(os.path.join(REFERENCE_DIR, "warp_params.p"), "rb"))
    return mod.LaneTracker(img_size=wp["image_width_height"], warped_size=wp["warped_width_height"],
                           cam_matrix=cal["cam_matrix"], dist_coeffs=cal["dist_coeffs"],
                           warp_matrices=(wp["M"], wp["Minv"]), mpp_conversion=(wp["mppv"], wp["mpph"]), **kw)


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)
