"""The oracle reproduces the live reference's recorded outputs (tests/golden/golden.json)."""
import warnings

import numpy as np
import pytest

import _fixtures as fx
from lane_tracker_b200 import synth
from oracle.tracker import OracleLaneTracker

GOLD = fx.golden()
CAL = synth.shipped_calibration()


def _state_matches(t, want, rel=1e-9):
    assert int(t.last_detection) == want["last_detection"]
    assert int(t.counter) == want["counter"] and int(t.success) == want["success"]
    assert bool(t.valid_lane_lines) == want["valid"]
    assert bool(t.detected_pixels) == want["detected_pixels"]
    for nm, key in (("last_left_coeffs", "last_left"), ("last_right_coeffs", "last_right"),
                    ("left_avg_coeffs", "left_avg"), ("right_avg_coeffs", "right_avg")):
        got = getattr(t, nm)
        if want[key] is None:
            assert got is None
        else:
            np.testing.assert_allclose(np.asarray(got), want[key], rtol=rel, atol=0)
    assert fx.avg_xy_digest(t) == want["avg_xy"]
    assert (None if t.average_curve_radius is None else int(t.average_curve_radius)) == want["radius"]
    assert [int(v) for v in t.average_curve_radii] == want["radii"]
    if want["ecc"] is None:
        assert t.eccentricity is None
    else:
        assert float(t.eccentricity) == pytest.approx(want["ecc"], rel=1e-12, abs=1e-15)
    if want["pix"] is not None:
        assert fx.pix_digest(t.left_y, t.left_x, t.right_y, t.right_x) == want["pix"]


@pytest.mark.parametrize("name", ["test1.jpg", "straight_lines1.jpg", "test5.jpg", "frame971.jpg"])
def test_bundled_frames_stage_digests(name):
    warnings.simplefilter("ignore")
    want = GOLD["images"][name]
    frame = fx.load_frame(name)
    assert fx.sha(frame) == want["frame"], "JPEG decode differs from the build container"
    t = OracleLaneTracker(**CAL)
    bv = t._remap(frame)
    assert fx.sha(t.trace["undistorted"][457:695]) == want["undistorted_rows_457_694"]
    assert fx.sha(bv) == want["bv"]
    m1 = t.filter_lane_points(bv, "bilateral", 15, 8, 35, 5, False, 65, 10, 140)
    m2 = t.filter_lane_points(bv, "neighborhood", 15, 5, 35, 5, False, 65, 10, 140)
    m3 = t.filter_lane_points(bv, "bilateral", 15, 8, 35, 5, True, 65, 10, 140)
    assert fx.sha(m1) == want["mask_bilateral"]
    assert fx.sha(m2) == want["mask_neighborhood"]
    assert fx.sha(m3) == want["mask_bilateral_noise"]
    for key, m, nsl in (("sws_bilateral", m1, 8), ("sws_neighborhood", m2, 50)):
        w = want[key]
        t.detected_pixels = False
        t.sliding_window_search(m, 30, 40, 20, 0.1, nsl, 0.25, 360, 30, 1.0)
        assert t.detected_pixels == w["detected"]
        if w["detected"]:
            assert fx.pix_digest(t.left_y, t.left_x, t.right_y, t.right_x) == w["pix"]
            assert [int(v) for v in t.left_window_centroids] == w["left_centroids"]
            assert [int(v) for v in t.right_window_centroids] == w["right_centroids"]
            lf, rf = t.fit_poly()
            np.testing.assert_allclose(lf, w["left_fit"], rtol=1e-9)
            np.testing.assert_allclose(rf, w["right_fit"], rtol=1e-9)
            t.check_validity(lf, rf)
            assert t.valid_lane_lines == w["valid"]


@pytest.mark.parametrize("name", ["test2.jpg", "straight_lines2.jpg"])
def test_bundled_frames_process(name):
    warnings.simplefilter("ignore")
    want = GOLD["images"][name]
    t = OracleLaneTracker(**CAL)
    out = t.process(fx.load_frame(name))
    assert fx.out_digest(out) == want["process_out"]
    assert fx.sha(out) == want["process_out_full"]          # including the putText overlays
    _state_matches(t, want["process_state"])


def test_scenario_sequence():
    """48 frames on one tracker: SWS -> band tracking -> 12-frame outage -> recovery."""
    warnings.simplefilter("ignore")
    sc = GOLD["scenario"]
    vid = synth.RoadVideo(sc["seed"])
    outage = fx.load_frame(sc["outage_frame"])
    t = OracleLaneTracker(**CAL, backend="cv2")  # cv2-backed ops: same restatement of the tracker logic, fast
    for rec in sc["frames"]:
        frame = outage if rec["kind"] == "outage" else vid.frame(rec["t"])
        assert fx.sha(frame) == rec["frame"], "synthetic generator is not deterministic"
        out = t.process(frame.copy())
        assert fx.out_digest(out) == rec["out"], rec["t"]
        assert fx.sha(out) == rec["out_full"], rec["t"]
        _state_matches(t, rec["state"])
    r = t.get_success_ratio()
    assert [float(r[0]), int(r[1]), int(r[2])] == sc["success_ratio"]


def test_scenario_prefix_numpy_backend():
    """Same scenario, pure-NumPy operators (no cv2 anywhere), first 4 frames."""
    warnings.simplefilter("ignore")
    sc = GOLD["scenario"]
    vid = synth.RoadVideo(sc["seed"])
    t = OracleLaneTracker(**CAL)
    for rec in sc["frames"][:4]:
        out = t.process(vid.frame(rec["t"]))
        assert fx.out_digest(out) == rec["out"], rec["t"]
        assert fx.sha(out) == rec["out_full"], rec["t"]      # text drawn from the glyph sprites
        _state_matches(t, rec["state"])


def _debug_view_frames(dv):
    vid = synth.RoadVideo(dv["seed"])
    outage = fx.load_frame(dv["outage_frame"])
    return [outage if kind == "outage" else vid.frame(t) for kind, t in dv["frames"]]


@pytest.mark.parametrize("mode", ["visualize_search", "split_view"])
def test_debug_views_golden(mode):
    """Reference debug views recorded by make_golden.py (lane_tracker.py:689-793, 1130-1209, utils.py:57-103)."""
    warnings.simplefilter("ignore")
    dv = GOLD["debug_views"]
    t = OracleLaneTracker(**CAL, backend="cv2")
    for i, (frame, rec) in enumerate(zip(_debug_view_frames(dv), dv[mode])):
        r = t.process(frame.copy(), **{mode: True})
        if mode == "visualize_search":
            assert fx.sha(r[0]) == rec["out"], i
            assert list(r[1].shape) == rec["vis_shape"] and fx.sha(r[1]) == rec["vis"], i
        else:
            assert list(r.shape) == rec["shape"] and fx.sha(r) == rec["canvas"], i


def test_debug_views_golden_numpy_backend():
    """Same, with the NumPy restatements of fillPoly / addWeighted / resize (first three frames: SWS view + band views)."""
    warnings.simplefilter("ignore")
    dv = GOLD["debug_views"]
    frames = _debug_view_frames(dv)[:3]
    t = OracleLaneTracker(**CAL)
    for i, frame in enumerate(frames):
        r = t.process(frame.copy(), visualize_search=True)
        assert fx.sha(r[1]) == dv["visualize_search"][i]["vis"], i
    t = OracleLaneTracker(**CAL)
    for i, frame in enumerate(frames[:2]):
        assert fx.sha(t.process(frame.copy(), split_view=True)) == dv["split_view"][i]["canvas"], i
