"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle and the golden digests
recorded from the reference.  Bit-exact for every integer/byte/index result; polynomial coefficients
within 1e-6 relative of np.polyfit (BASELINE.json north_star), in practice ~1e-11."""
import warnings

import numpy as np
import pytest

import _fixtures as fx
from lane_tracker_b200 import synth
from oracle import cvops
from oracle.tracker import OracleLaneTracker, ATTEMPT2

pytestmark = pytest.mark.gpu

CAL = synth.shipped_calibration()
GOLD = fx.golden()
FIT_RTOL = 1e-6   # tolerance stated by north_star for coefficients vs np.polyfit


@pytest.fixture(scope="module")
def torch_mod():
    import torch
    assert torch.cuda.is_available()
    return torch


@pytest.fixture(scope="module")
def bt(torch_mod):
    from lane_tracker_b200 import BatchedLaneTracker
    t = BatchedLaneTracker(4, **CAL)
    yield t
    t.close()


@pytest.fixture(scope="module")
def frames_np():
    rng = np.random.default_rng(11)
    return np.stack([fx.load_frame("test1.jpg"), synth.RoadVideo(0).frame(3), fx.load_frame("straight_lines2.jpg"),
                     rng.integers(0, 256, (720, 1280, 3), dtype=np.uint8)])


@pytest.fixture(scope="module")
def oracle_stages(frames_np):
    """Oracle intermediates for the four fixture frames (NumPy restatements only)."""
    out = []
    for f in frames_np:
        o = OracleLaneTracker(**CAL)
        bv = o._remap(f)
        rec = dict(bv=bv, und=o.trace["undistorted"])
        m1 = o.filter_lane_points(bv, "bilateral", 15, 8, 35, 5, False, 65, 10, 140)
        rec.update(mask1=m1, r_plane=o.trace["r_plane"], b_plane=o.trace["b_plane"], r_tophat=o.trace["r_tophat"],
                   b_tophat=o.trace["b_tophat"], merged1=o.trace["merged"])
        rec["mask2"] = o.filter_lane_points(bv, "neighborhood", 15, 5, 35, 5, False, 65, 10, 140)
        rec["mask3"] = o.filter_lane_points(bv, "bilateral", 15, 8, 35, 5, True, 65, 10, 140)
        out.append(rec)
    return out


def _mism(a, b):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, (a.shape, b.shape)
    return int((a != b).sum())


def test_library_loaded_is_in_tree():
    from lane_tracker_b200 import _lib
    lib = _lib.load()
    assert "lane_tracker_b200/liblane_tracker_b200.so" in _lib.LIB_PATH and lib.lt_abi_version() == _lib.LT_ABI_VERSION


def test_coordinate_tables(bt):
    U, V = cvops.undistort_map_q5(CAL["cam_matrix"], CAL["dist_coeffs"], 1280, 720)
    got = bt.debug_read("undistort_map")
    assert _mism(got[..., 0], U) == 0 and _mism(got[..., 1], V) == 0
    X, Y = cvops.perspective_map_q5(CAL["warp_matrices"][0], 1080, 1100)
    got = bt.debug_read("bv_map")
    assert _mism(got[..., 0], X) == 0 and _mism(got[..., 1], Y) == 0
    X, Y = cvops.perspective_map_q5(CAL["warp_matrices"][1], 1280, 720)
    got = bt.debug_read("overlay_map")
    assert _mism(got[..., 0], X) == 0 and _mism(got[..., 1], Y) == 0
    # the materialised undistorted rows cover exactly what the bird's-eye view samples (SURVEY: rows 457-694)
    assert bt.geometry["roi_rows"] == (457, 695)


def test_remap_bit_exact(bt, torch_mod, frames_np, oracle_stages):
    d = torch_mod.as_tensor(frames_np).cuda()
    bv = bt.remap(d).cpu().numpy()
    for i, rec in enumerate(oracle_stages):
        assert _mism(bv[i], rec["bv"]) == 0, i
        assert _mism(bt.debug_read("r_plane", i), rec["r_plane"]) == 0, i
        assert _mism(bt.debug_read("b_plane", i), rec["b_plane"]) == 0, i
    assert fx.sha(bv[0]) == GOLD["images"]["test1.jpg"]["bv"]
    assert fx.sha(bv[2]) == GOLD["images"]["straight_lines2.jpg"]["bv"]


def test_filter_masks_bit_exact(bt, torch_mod, oracle_stages):
    bvs = torch_mod.as_tensor(np.stack([r["bv"] for r in oracle_stages])).cuda()
    m1 = bt.filter_lane_points(bvs, "bilateral", 15, 8, 35, 5, False, 65, 10, 140).cpu().numpy()
    for i, rec in enumerate(oracle_stages):
        assert _mism(bt.debug_read("r_tophat", i), rec["r_tophat"]) == 0, i
        assert _mism(bt.debug_read("b_tophat", i), rec["b_tophat"]) == 0, i
        assert _mism(bt.debug_read("merged", i), rec["merged1"]) == 0, i
        assert _mism(m1[i], rec["mask1"]) == 0, i
    m2 = bt.filter_lane_points(bvs, "neighborhood", 15, 5, 35, 5, False, 65, 10, 140).cpu().numpy()
    m3 = bt.filter_lane_points(bvs, "bilateral", 15, 8, 35, 5, True, 65, 10, 140).cpu().numpy()
    for i, rec in enumerate(oracle_stages):
        assert _mism(m2[i], rec["mask2"]) == 0, i
        assert _mism(m3[i], rec["mask3"]) == 0, i
    g = GOLD["images"]["test1.jpg"]
    assert fx.sha(m1[0]) == g["mask_bilateral"] and fx.sha(m2[0]) == g["mask_neighborhood"]
    assert fx.sha(m3[0]) == g["mask_bilateral_noise"]


@pytest.mark.parametrize("k,C", [(25, 8), (7, 2), (64, 3)])
def test_filter_other_kernel_sizes(bt, torch_mod, oracle_stages, k, C):
    bv = oracle_stages[0]["bv"]
    o = OracleLaneTracker(**CAL)
    want = o.filter_lane_points(bv, "bilateral", k, C, k + 6, C, False, 65, 10, 140)
    got = bt.filter_lane_points(torch_mod.as_tensor(bv[None]).cuda(), "bilateral", k, C, k + 6, C).cpu().numpy()[0]
    assert _mism(got, want) == 0
    kb = k | 1
    want = o.filter_lane_points(bv, "neighborhood", kb, C, kb + 8, C + 1, False, 65, 10, 140)
    got = bt.filter_lane_points(torch_mod.as_tensor(bv[None]).cuda(), "neighborhood", kb, C, kb + 8, C + 1).cpu().numpy()[0]
    assert _mism(got, want) == 0


def _oracle_sws(mask, nsl, partial=1.0, **kw):
    o = OracleLaneTracker(**CAL)
    args = dict(window_width=30, window_height=40, search_range=20, mu=0.1, no_success_limit=nsl,
                start_slice=0.25, ignore_sides=360, ignore_bottom=30, partial=partial)
    args.update(kw)
    o.sliding_window_search(mask, **args)
    return o


def test_sliding_window_search_pixel_sets(bt, torch_mod, oracle_stages):
    cases = []
    for rec in oracle_stages:
        cases += [(rec["mask1"], 8, 1.0), (rec["mask2"], 50, 1.0), (rec["mask1"], 8, 0.5)]
    sparse = np.zeros((1100, 1080), np.uint8)                 # edge cases: nearly empty / one-sided masks
    sparse[1040:1060, 400:404] = 255
    cases += [(sparse, 8, 1.0), (np.zeros((1100, 1080), np.uint8), 8, 1.0)]
    for ci, (mask, nsl, partial) in enumerate(cases):
        o = _oracle_sws(mask, nsl, partial)
        px, cents, det = bt.sliding_window_search(torch_mod.as_tensor(mask[None]).cuda(), 30, 40, 20, 0.1, nsl,
                                                  0.25, 360, 30, partial)
        assert bool(det[0]) == o.detected_pixels, ci
        got_c = cents[0]
        assert got_c[0] == o.trace["sws_centroids"][0] and got_c[1] == o.trace["sws_centroids"][1], ci
        if o.detected_pixels:
            (ly, lx), (ry, rx) = px[0]
            assert np.array_equal(ly, o.left_y) and np.array_equal(lx, o.left_x), ci
            assert np.array_equal(ry, o.right_y) and np.array_equal(rx, o.right_x), ci


def test_sliding_window_other_parameters(bt, torch_mod, oracle_stages):
    mask = oracle_stages[1]["mask1"]
    for kw in (dict(window_width=40, window_height=55, search_range=35, mu=0.3), dict(ignore_sides=200, start_slice=0.5),
               dict(window_width=17, window_height=25, ignore_bottom=0)):
        o = _oracle_sws(mask, 8, 1.0, **kw)
        a = dict(window_width=30, window_height=40, search_range=20, mu=0.1, no_success_limit=8, start_slice=0.25,
                 ignore_sides=360, ignore_bottom=30, partial=1.0)
        a.update(kw)
        px, cents, det = bt.sliding_window_search(torch_mod.as_tensor(mask[None]).cuda(), a["window_width"],
                                                  a["window_height"], a["search_range"], a["mu"], a["no_success_limit"],
                                                  a["start_slice"], a["ignore_sides"], a["ignore_bottom"], a["partial"])
        assert bool(det[0]) == o.detected_pixels
        assert cents[0][0] == o.trace["sws_centroids"][0] and cents[0][1] == o.trace["sws_centroids"][1], kw
        if o.detected_pixels:
            (ly, lx), (ry, rx) = px[0]
            assert np.array_equal(ly, o.left_y) and np.array_equal(lx, o.left_x), kw
            assert np.array_equal(ry, o.right_y) and np.array_equal(rx, o.right_x), kw


def test_band_search_and_fit(bt, torch_mod, oracle_stages):
    warnings.simplefilter("ignore")
    mask = oracle_stages[1]["mask1"]
    o = _oracle_sws(mask, 8)
    lf, rf = o.fit_poly()
    for bw, partial, ib in ((25, 1.0, 30), (30, 0.5, 30), (40, 1.0, 0)):
        o.last_left_coeffs, o.last_right_coeffs = lf, rf
        o.band_search(mask, bw, ib, partial)
        px, det = bt.band_search(torch_mod.as_tensor(mask[None]).cuda(), np.stack([lf, rf])[None], bw, ib, partial)
        assert bool(det[0]) == o.detected_pixels
        (ly, lx), (ry, rx) = px[0]
        assert np.array_equal(ly, o.left_y) and np.array_equal(lx, o.left_x)
        assert np.array_equal(ry, o.right_y) and np.array_equal(rx, o.right_x)
        fits = bt.fit_poly([[(ly, lx), (ry, rx)]])[0]
        wl, wr = o.fit_poly()
        np.testing.assert_allclose(fits[0], wl, rtol=FIT_RTOL, atol=0)
        np.testing.assert_allclose(fits[1], wr, rtol=FIT_RTOL, atol=0)
        assert np.max(np.abs(fits[0] / wl - 1)) < 1e-9 and np.max(np.abs(fits[1] / wr - 1)) < 1e-9


def test_fit_poly_rank_deficient_and_random(bt):
    warnings.simplefilter("ignore")
    rng = np.random.default_rng(3)
    sets = []
    ys = np.full(20, 700); xs = rng.integers(300, 340, 20)            # one row
    sets.append([(ys, xs), (np.array([10, 10, 900]), np.array([5, 9, 700]))])   # two rows
    for _ in range(6):
        n = int(rng.integers(3, 4000))
        y = rng.integers(0, 1100, n); x = np.clip((2e-4 * (y - 500.0) ** 2 + 0.1 * y + 300 + rng.normal(0, 4, n)), 0, 1079).astype(int)
        y2 = rng.integers(0, 1100, n); x2 = rng.integers(0, 1080, n)
        sets.append([(y, x), (y2, x2)])
    fits = bt.fit_poly(sets)
    for s, ps in enumerate(sets):
        for side in range(2):
            want = np.polyfit(ps[side][0], ps[side][1], 2)
            np.testing.assert_allclose(fits[s, side], want, rtol=FIT_RTOL, atol=1e-9 * np.abs(want).max())


def test_check_validity_and_poly_points(bt):
    rng = np.random.default_rng(5)
    o = OracleLaneTracker(**CAL)
    fits = []
    for t in range(300):
        a = rng.normal(0, 2e-4); b = rng.normal(0, 0.3); c = rng.uniform(200, 600)
        sep = rng.uniform(60, 260)
        fits.append([[a, b - 2 * a * 1099, a * 1099 ** 2 - b * 1099 + c],
                     [a + rng.normal(0, 5e-5), b + rng.normal(0, 0.1) - 2 * a * 1099, a * 1099 ** 2 - b * 1099 + c + sep]])
    fits = np.array(fits)
    valid, diffs = bt.check_validity(fits)
    assert 10 < valid.sum() < 290
    for i, f in enumerate(fits):
        o.check_validity(f[0], f[1])
        assert bool(valid[i]) == o.valid_lane_lines, i
        assert tuple(diffs[i]) == tuple(o.trace["validity"]), i
    for partial in (1.0, 0.5, 0.73):
        xs, cnt = bt.get_poly_points(fits[:40], partial)
        xs, cnt = xs.cpu().numpy(), cnt.cpu().numpy()
        for i, f in enumerate(fits[:40]):
            ly, lx, ry, rx = o.get_poly_points(f[0], f[1], partial)
            assert cnt[i, 0] == len(lx) and cnt[i, 1] == len(rx)
            assert np.array_equal(xs[i, 0, :len(lx)], lx) and np.array_equal(xs[i, 1, :len(rx)], rx)


def test_draw_lane_overlay(bt, torch_mod, frames_np):
    rng = np.random.default_rng(9)
    o = OracleLaneTracker(**CAL, render_text=False)      # lt_draw_lane is the polygon + un-warp + blend stage only
    for t in range(6):
        a = rng.normal(0, 3e-4) * (4 if t % 2 else 1); b = rng.normal(0, 0.4); c = rng.uniform(150, 600)
        lf = np.array([a, b - 2 * a * 1099, a * 1099 ** 2 - b * 1099 + c])
        rf = lf + np.array([rng.normal(0, 5e-5), rng.normal(0, 0.05), rng.uniform(120, 260)])
        partial = 1.0 if t < 4 else 0.5
        o.left_avg_y, o.left_avg_x, o.right_avg_y, o.right_avg_x = o.get_poly_points(lf, rf, partial)
        if len(o.left_avg_x) == 0 or len(o.right_avg_x) == 0:
            continue
        want = o.draw_lane(frames_np[t % 4])
        xs, cnt = bt.get_poly_points(np.stack([lf, rf])[None], partial)
        got = bt.draw_lane(torch_mod.as_tensor(frames_np[t % 4][None]).cuda(), xs, cnt).cpu().numpy()[0]
        rows = bt.debug_read("draw_lane_rows", 0)
        lo, hi = o.trace["lane_rows"]
        filled = hi >= lo
        assert np.array_equal(rows[filled, 0], lo[filled]) and np.array_equal(rows[filled, 1], hi[filled]), t
        assert np.all(rows[~filled, 1] < rows[~filled, 0]), t
        assert _mism(got, want) == 0, t


def _check_result_against_golden(res, want, st, lx, rx):
    assert int(res["last_detection"]) == want["last_detection"]
    assert int(res["counter"]) == want["counter"] and int(res["success"]) == want["success"]
    assert bool(res["valid_lane_lines"]) == want["valid"]
    assert bool(res["detected_pixels"]) == want["detected_pixels"]
    if want["last_left"] is not None:
        np.testing.assert_allclose(np.array(st.last_left[:]), want["last_left"], rtol=FIT_RTOL)
        np.testing.assert_allclose(np.array(st.last_right[:]), want["last_right"], rtol=FIT_RTOL)
        np.testing.assert_allclose(np.array(st.left_avg[:]), want["left_avg"], rtol=FIT_RTOL)
        np.testing.assert_allclose(np.array(st.right_avg[:]), want["right_avg"], rtol=FIT_RTOL)
    assert st.n_left_avg == want["n_left_avg"] and st.n_right_avg == want["n_right_avg"]
    bh = 1100
    digest = fx.sha(np.concatenate([np.arange(bh - len(lx), bh), lx.astype(np.int64), np.arange(bh - len(rx), bh),
                                    rx.astype(np.int64)]).astype(np.int64)) if want["n_left_avg"] else None
    if digest is not None:
        assert digest == want["avg_xy"]
    if want["radius"] is not None:
        assert int(st.average_curve_radius) == want["radius"]
        assert [int(st.radii[i]) for i in range(st.radii_len)] == want["radii"]
        assert float(st.eccentricity) == pytest.approx(want["ecc"], rel=1e-12, abs=1e-15)


def test_process_scenario_matches_reference_golden(torch_mod):
    """48 frames through the drop-in class: SWS -> band tracking -> 12-frame outage -> recovery."""
    from lane_tracker_b200 import LaneTracker
    sc = GOLD["scenario"]
    vid = synth.RoadVideo(sc["seed"])
    outage = fx.load_frame(sc["outage_frame"])
    lt = LaneTracker(**CAL)
    for rec in sc["frames"]:
        frame = outage if rec["kind"] == "outage" else vid.frame(rec["t"])
        keep = frame.copy()
        out = lt.process(frame)
        assert np.array_equal(frame, keep)
        w = rec["state"]
        st, lx, rx = lt._bt.get_state(0)
        _check_result_against_golden(lt.last_result, w, st, lx, rx)
        assert fx.out_digest(out) == rec["out"], rec["t"]
        assert fx.sha(out) == rec["out_full"], rec["t"]      # whole frame, putText overlays included
        if w["pix"] is not None:
            assert fx.pix_digest(lt.left_y, lt.left_x, lt.right_y, lt.right_x) == w["pix"], rec["t"]
    r = lt.get_success_ratio()
    assert [float(r[0]), int(r[1]), int(r[2])] == sc["success_ratio"]


def test_process_bundled_frames_match_reference_golden(torch_mod):
    """All 11 bundled frames (always two attempts + sliding-window search, all invalid), batched."""
    from lane_tracker_b200 import BatchedLaneTracker
    names = fx.frame_names()
    frames = np.stack([fx.load_frame(n) for n in names])
    bt11 = BatchedLaneTracker(len(names), **CAL)
    bt11.set_capture(True)
    d = torch_mod.as_tensor(frames).cuda()
    out = torch_mod.empty_like(d)
    res = bt11.process(d, out)
    out = out.cpu().numpy()
    for i, n in enumerate(names):
        g = GOLD["images"][n]
        assert fx.sha(frames[i]) == g["frame"]
        assert int(res[i]["attempts"]) == 2 and not res[i]["valid_lane_lines"]
        assert fx.out_digest(out[i]) == g["process_out"], n
        assert fx.sha(out[i]) == g["process_out_full"], n
        assert fx.sha(bt11.debug_read("mask", i)) == g["mask_neighborhood"], n
        for attempt, key in ((0, "sws_bilateral"), (1, "sws_neighborhood")):
            w = g[key]
            sides, cents = bt11.read_capture(i, attempt)
            det = len(sides[0][0]) > 0 and len(sides[1][0]) > 0
            assert det == w["detected"], (n, key)
            if det:
                assert fx.pix_digest(sides[0][0], sides[0][1], sides[1][0], sides[1][1]) == w["pix"], (n, key)
                assert cents[0] == w["left_centroids"] and cents[1] == w["right_centroids"], (n, key)
        w2 = g["sws_neighborhood"]
        if w2["detected"]:
            np.testing.assert_allclose(res[i]["left_fit"], w2["left_fit"], rtol=FIT_RTOL)
            np.testing.assert_allclose(res[i]["right_fit"], w2["right_fit"], rtol=FIT_RTOL)
        w1 = g["sws_bilateral"]
        if w1["detected"]:
            np.testing.assert_allclose(res[i]["first_left_fit"], w1["left_fit"], rtol=FIT_RTOL)
            np.testing.assert_allclose(res[i]["first_right_fit"], w1["right_fit"], rtol=FIT_RTOL)
            assert bool(res[i]["first_valid"]) == w1["valid"]
    bt11.close()


def test_batched_streams_are_independent(torch_mod):
    """Stream s of a batch behaves exactly like a single-stream tracker fed the same frames."""
    from lane_tracker_b200 import BatchedLaneTracker
    S, T = 3, 5
    vids = [synth.RoadVideo(s) for s in range(S)]
    b = BatchedLaneTracker(S, **CAL)
    singles = [BatchedLaneTracker(1, **CAL) for _ in range(S)]
    for t in range(T):
        fr = np.stack([v.frame(t) for v in vids])
        d = torch_mod.as_tensor(fr).cuda()
        out = torch_mod.empty_like(d)
        res = b.process(d, out)
        for s in range(S):
            o1 = torch_mod.empty_like(d[s:s + 1])
            r1 = singles[s].process(d[s:s + 1].contiguous(), o1)
            assert r1[0].tobytes() == res[s].tobytes()
            assert torch_mod.equal(o1[0], out[s])
    for t in singles + [b]:
        t.close()


def test_state_roundtrip_and_teacher_forcing(torch_mod):
    from lane_tracker_b200 import BatchedLaneTracker
    vid = synth.RoadVideo(2)
    a = BatchedLaneTracker(1, **CAL)
    b = BatchedLaneTracker(1, **CAL)
    for t in range(3):
        a.process(torch_mod.as_tensor(vid.frame(t)[None]).cuda())
    st, lx, rx = a.get_state(0)
    b.set_state(0, st, lx, rx)
    f = torch_mod.as_tensor(vid.frame(3)[None]).cuda()
    oa, ob = torch_mod.empty_like(f), torch_mod.empty_like(f)
    ra, rb = a.process(f, oa), b.process(f, ob)
    assert ra[0].tobytes() == rb[0].tobytes() and torch_mod.equal(oa, ob)
    a.reset()
    assert a.get_state(0)[0].last_detection == 5 and a.get_state(0)[0].counter == 0
    a.close(); b.close()


def test_error_conventions(bt, torch_mod):
    from lane_tracker_b200 import LaneTracker, _lib
    lt = LaneTracker(**CAL)
    with pytest.raises(ValueError, match="Unexpected filter mode"):
        lt.filter_lane_points(np.zeros((1100, 1080, 3), np.uint8), filter_type="gaussian")
    with pytest.raises(ValueError):
        lt.process(np.zeros((10, 10, 3), np.uint8))
    with pytest.raises(_lib.LaneTrackerError):
        bt.filter_lane_points(torch_mod.zeros((1, 1100, 1080, 3), dtype=torch_mod.uint8, device="cuda"), "neighborhood", 14, 5, 35, 5)


def test_host_pipeline_matches_sequential_process(torch_mod):
    """The overlapped host-to-host pipeline returns exactly what sequential process() calls return."""
    from lane_tracker_b200 import BatchedLaneTracker, HostPipeline
    S, T = 2, 7
    vids = [synth.RoadVideo(10 + s) for s in range(S)]
    batches = [torch_mod.from_numpy(np.stack([v.frame(t) for v in vids])).pin_memory() for t in range(T)]
    a = BatchedLaneTracker(S, **CAL)
    b = BatchedLaneTracker(S, **CAL)
    want = []
    for t in range(T):
        d = batches[t].cuda()
        out = torch_mod.empty_like(d)
        res = a.process(d, out)
        want.append((out.cpu().numpy(), res.copy()))
    pipe = HostPipeline(b, depth=3)
    got = []
    for t in range(T):
        pipe.submit(batches[t])
        for out, res in pipe.ready():
            got.append((out.numpy().copy(), res.copy()))
    for out, res in pipe.drain():
        got.append((out.numpy().copy(), res.copy()))
    assert len(got) == T
    for t in range(T):
        assert np.array_equal(got[t][0], want[t][0]), t
        assert got[t][1].tobytes() == want[t][1].tobytes(), t
    a.close(); b.close()


@pytest.mark.parametrize("scale", [1.5, 3.0])
def test_scaled_geometry_matches_oracle(torch_mod, scale):
    """BASELINE.json config 5: 1920x1080 / 3840x2160 frames with rescaled calibration.  The
    reference's pixel-valued constants do not scale, so every frame takes both attempts + sliding-window search."""
    warnings.simplefilter("ignore")
    from lane_tracker_b200 import BatchedLaneTracker
    cal = synth.shipped_calibration(scale)
    vid = synth.RoadVideo(4, scale=scale)
    frame = vid.frame(2)
    o = OracleLaneTracker(**cal, backend="cv2")
    want = o.process(frame.copy(), n_tries=2)
    b = BatchedLaneTracker(1, **cal)
    b.set_capture(True)
    d = torch_mod.as_tensor(frame[None]).cuda()
    out = torch_mod.empty_like(d)
    res = b.process(d, out)[0]
    att = o.trace["attempts"]
    assert int(res["attempts"]) == len(att)
    assert _mism(b.debug_read("mask", 0), att[-1]["mask"]) == 0
    assert bool(res["detected_pixels"]) == att[-1]["detected"] and bool(res["valid_lane_lines"]) == att[-1]["valid"]
    if att[-1]["detected"]:
        sides, _ = b.read_capture(0, len(att) - 1)
        assert np.array_equal(sides[0][0], att[-1]["left_y"]) and np.array_equal(sides[0][1], att[-1]["left_x"])
        assert np.array_equal(sides[1][0], att[-1]["right_y"]) and np.array_equal(sides[1][1], att[-1]["right_x"])
        np.testing.assert_allclose(res["left_fit"], att[-1]["left_fit"], rtol=FIT_RTOL)
        np.testing.assert_allclose(res["right_fit"], att[-1]["right_fit"], rtol=FIT_RTOL)
    assert _mism(out.cpu().numpy()[0], want) == 0
    # first attempt ('bilateral' filter) mask as well
    m1 = b.filter_lane_points(None, "bilateral", 15, 8, 35, 5).cpu().numpy()[0]
    assert _mism(m1, att[0]["mask"]) == 0
    b.close()


def test_sliding_window_adversarial_masks(bt, torch_mod):
    import _masks
    masks = _masks.random_masks(36, seed=5)
    for i in range(0, len(masks), 4):
        batch = np.stack(masks[i:i + 4])
        for nsl, partial, mu in ((8, 1.0, 0.1), (50, 1.0, 0.1), (8, 0.5, 0.35)):
            px, cents, det = bt.sliding_window_search(torch_mod.as_tensor(batch).cuda(), 30, 40, 20, mu, nsl, 0.25, 360,
                                                      30, partial)
            for j in range(batch.shape[0]):
                o = _oracle_sws(batch[j], nsl, partial, mu=mu)
                assert bool(det[j]) == o.detected_pixels, (i + j, nsl)
                assert cents[j][0] == o.trace["sws_centroids"][0] and cents[j][1] == o.trace["sws_centroids"][1], (i + j, nsl)
                if o.detected_pixels:
                    (ly, lx), (ry, rx) = px[j]
                    assert np.array_equal(ly, o.left_y) and np.array_equal(lx, o.left_x), (i + j, nsl)
                    assert np.array_equal(ry, o.right_y) and np.array_equal(rx, o.right_x), (i + j, nsl)


def test_demo1_preset_on_bundled_frames(torch_mod):
    """Extra coverage (clearly labelled, not the headline parity): with the Demo-1 windows of
    tracker_settings.md:28-34 nine of the eleven bundled frames validate, so real photographs reach band search,
    mask_noise and the success branch.  Three passes of every frame, all 11 batched as independent streams."""
    warnings.simplefilter("ignore")
    from lane_tracker_b200 import BatchedLaneTracker, presets
    names = fx.frame_names()
    frames = np.stack([fx.load_frame(n) for n in names])
    V = presets.DEMO_1["validity"]
    oracles = []
    for _ in names:
        o = OracleLaneTracker(**CAL, backend="cv2")
        o.validity = dict(min_d1=V["min_dist_y1"], max_d1=V["max_dist_y1"], min_d2=V["min_dist_y2"],
                          max_d2=V["max_dist_y2"], min_d3=V["min_dist_y3"], max_d3=V["max_dist_y3"],
                          tan=V["tangent_thresh"])
        oracles.append(o)
    b = BatchedLaneTracker(len(names), **CAL)
    b.set_validity(**V)
    b.set_capture(True)
    d = torch_mod.as_tensor(frames).cuda()
    n_valid = n_band = 0
    for it in range(3):
        out = torch_mod.empty_like(d)
        res = b.process(d, out, **presets.DEMO_1["process"])
        out = out.cpu().numpy()
        for i, o in enumerate(oracles):
            want = o.process(frames[i].copy(), **presets.DEMO_1["process"])
            att = o.trace["attempts"]
            assert int(res[i]["attempts"]) == len(att), (names[i], it)
            assert bool(res[i]["valid_lane_lines"]) == o.valid_lane_lines, (names[i], it)
            assert int(res[i]["search_mode"]) == (1 if att[-1]["mode"] == "bs" else 0)
            assert int(res[i]["last_detection"]) == o.last_detection
            assert _mism(out[i], want) == 0, (names[i], it)
            if att[-1]["detected"]:
                sides, _ = b.read_capture(i, len(att) - 1)
                assert np.array_equal(sides[0][1], att[-1]["left_x"]) and np.array_equal(sides[1][0], att[-1]["right_y"])
                np.testing.assert_allclose(res[i]["left_fit"], att[-1]["left_fit"], rtol=FIT_RTOL)
                np.testing.assert_allclose(res[i]["right_fit"], att[-1]["right_fit"], rtol=FIT_RTOL)
            if o.valid_lane_lines:
                n_valid += 1
                n_band += att[-1]["mode"] == "bs"
                assert int(res[i]["average_curve_radius"]) == o.average_curve_radius
                assert float(res[i]["eccentricity"]) == pytest.approx(o.eccentricity, rel=1e-12, abs=1e-15)
    assert n_valid == 27 and n_band == 18
    b.set_validity()            # back to the shipped constants: nothing validates any more
    b.reset()
    res = b.process(d, None)
    assert not res["valid_lane_lines"].any()
    b.close()


@pytest.mark.parametrize("case", range(6))
def test_process_randomised_options_and_state_machine(torch_mod, case):
    """Constructor options (n_fail, n_reset, n_average) and process() keyword options drawn at random; a short
    sequence with an outage; every frame compared with the oracle (outputs, decisions, pixel sets, state)."""
    warnings.simplefilter("ignore")
    from lane_tracker_b200 import LaneTracker
    rng = np.random.default_rng(100 + case)
    ctor = dict(n_fail=int(rng.integers(1, 6)), n_reset=int(rng.integers(0, 4)), n_average=int(rng.integers(1, 6)))
    kw = dict(ksize_r=int(rng.integers(8, 25)), C_r=int(rng.integers(3, 10)), ksize_b=int(rng.integers(20, 45)),
              C_b=int(rng.integers(3, 8)), mask_noise=bool(rng.integers(0, 2)), noise_thresh=int(rng.integers(130, 150)),
              window_width=int(rng.choice([24, 30, 36])), window_height=int(rng.choice([30, 40, 50])),
              search_range=int(rng.integers(12, 30)), mu=float(rng.choice([0.0, 0.1, 0.3])),
              no_success_limit=int(rng.integers(3, 12)), bandwidth=int(rng.integers(15, 40)),
              partial=float(rng.choice([1.0, 0.5, 0.8])), n_tries=int(rng.choice([1, 2])))
    vid = synth.RoadVideo(20 + case)
    blank = np.full((720, 1280, 3), 90, np.uint8)
    plan = [0, 1, 2, 3, "x", "x", "x", "x", "x", "x", 4, 5, 6] if case % 2 == 0 else [0, "x", 1, 2, "x", "x", 3, 4, 5]
    gpu = LaneTracker(**CAL, **ctor)
    ref = OracleLaneTracker(**CAL, **ctor, backend="cv2")
    for step, item in enumerate(plan):
        frame = blank if item == "x" else vid.frame(item)
        out = gpu.process(frame, **kw)
        want = ref.process(frame.copy(), **kw)
        tag = (case, step, ctor, kw)
        assert gpu.valid_lane_lines == ref.valid_lane_lines, tag
        assert gpu.detected_pixels == ref.detected_pixels, tag
        assert gpu.last_detection == ref.last_detection and gpu.success == ref.success, tag
        assert _mism(out, want) == 0, tag
        assert gpu.average_curve_radii == ref.average_curve_radii, tag
        assert len(gpu.left_fit_coeffs) == len(ref.left_fit_coeffs), tag
        for a, b in zip(gpu.left_fit_coeffs + gpu.right_fit_coeffs, ref.left_fit_coeffs + ref.right_fit_coeffs):
            assert a.size == b.size, tag
            if a.size:
                np.testing.assert_allclose(a, b, rtol=FIT_RTOL)
        if ref.left_avg_coeffs is not None:
            np.testing.assert_allclose(gpu.left_avg_coeffs, ref.left_avg_coeffs, rtol=FIT_RTOL)
            assert np.array_equal(gpu.left_avg_x, ref.left_avg_x) and np.array_equal(gpu.right_avg_x, ref.right_avg_x), tag
            assert np.array_equal(gpu.left_avg_y, ref.left_avg_y), tag
            assert gpu.average_curve_radius == ref.average_curve_radius, tag
        if ref.left_x is not None and ref.detected_pixels:
            assert np.array_equal(gpu.left_x, ref.left_x) and np.array_equal(gpu.right_y, ref.right_y), tag


def test_inplace_annotation_and_roi_pipeline(torch_mod):
    """d_out == d_frames rewrites only the rows the overlay can reach; HostPipeline(inplace=True) moves only the
    rows the tracker reads / the overlay changes across PCIe and leaves the same image in the caller's buffer."""
    from lane_tracker_b200 import BatchedLaneTracker, HostPipeline
    S, T = 2, 6
    vids = [synth.RoadVideo(30 + s) for s in range(S)]
    frames = [np.stack([v.frame(t) for v in vids]) for t in range(T)]
    a = BatchedLaneTracker(S, **CAL)
    b = BatchedLaneTracker(S, **CAL)
    c = BatchedLaneTracker(S, **CAL)
    g = a.geometry
    assert 0 <= g["source_rows"][0] < g["source_rows"][1] <= 720
    assert 0 < g["overlay_rows"][0] < g["overlay_rows"][1] <= 720
    want = []
    for t in range(T):
        d = torch_mod.as_tensor(frames[t]).cuda()
        out = torch_mod.empty_like(d)
        res = a.process(d, out)
        want.append((out.cpu().numpy(), res.copy()))
        d2 = d.clone()
        res2 = b.process(d2, d2)                       # in place on the device
        assert np.array_equal(d2.cpu().numpy(), want[-1][0]) and res2.tobytes() == res.tobytes(), t
    assert (want[-1][0] != frames[-1]).any()           # the overlay really changed pixels
    pipe = HostPipeline(c, depth=3, inplace=True)
    host = [torch_mod.from_numpy(f.copy()).pin_memory() for f in frames]
    got = []
    for t in range(T):
        pipe.submit(host[t])
        for fr, res in pipe.ready():
            got.append((fr, res.copy()))
    for fr, res in pipe.drain():
        got.append((fr, res.copy()))
    assert len(got) == T
    for t in range(T):
        assert got[t][0] is host[t]
        assert np.array_equal(host[t].numpy(), want[t][0]), t
        assert got[t][1].tobytes() == want[t][1].tobytes(), t
    for x in (a, b, c):
        x.close()


def test_fused_remap_variant(torch_mod):
    """North-star item (1): the fused single-resample remap, reported separately under a stated tolerance.
    It is deterministic (equals its NumPy restatement bit for bit) and, against the exact two-stage pipeline on the
    reference's 11 bundled frames: mask IoU >= 0.6 per frame and >= 0.8 on average."""
    from lane_tracker_b200 import BatchedLaneTracker
    names = fx.frame_names()
    frames = np.stack([fx.load_frame(n) for n in names])
    b = BatchedLaneTracker(len(names), **CAL)
    d = torch_mod.as_tensor(frames).cuda()
    exact_bv = b.remap(d).cpu().numpy()
    exact_mask = b.filter_lane_points(None, "bilateral", 15, 8, 35, 5).cpu().numpy() > 0
    b.set_remap_mode("fused")
    fused_bv = b.remap(d).cpu().numpy()
    fused_mask = b.filter_lane_points(None, "bilateral", 15, 8, 35, 5).cpu().numpy() > 0
    for i in (0, 5):
        want = cvops.fused_bird_view(frames[i], CAL["cam_matrix"], CAL["dist_coeffs"], CAL["warp_matrices"][0], 1080, 1100)
        assert _mism(fused_bv[i], want) == 0, names[i]
    ious, changed = [], []
    for i in range(len(names)):
        inter = (exact_mask[i] & fused_mask[i]).sum()
        union = (exact_mask[i] | fused_mask[i]).sum()
        ious.append(inter / max(union, 1))
        changed.append(float((exact_bv[i] != fused_bv[i]).any(axis=2).mean()))
    print("fused-vs-exact mask IoU per frame:", [round(v, 3) for v in ious], "mean", round(float(np.mean(ious)), 3),
          "| changed BV pixels:", round(float(np.mean(changed)), 3))
    assert min(ious) >= 0.6 and np.mean(ious) >= 0.8
    assert 0.05 < np.mean(changed) < 0.9            # it really is a different resampling
    # the tracker runs end to end in this mode
    res = b.process(d, None)
    assert (res["attempts"] == 2).all()
    b.set_remap_mode("exact")
    assert _mism(b.remap(d).cpu().numpy(), exact_bv) == 0
    b.close()


def test_abi_error_paths_partial_batches_and_lifetime(torch_mod):
    import ctypes as C
    from lane_tracker_b200 import BatchedLaneTracker, _lib
    lib = _lib.load()
    # constructor validation (no exceptions cross the ABI: negative return + message)
    bad = dict(CAL, img_size=(1281, 720))
    with pytest.raises(_lib.LaneTrackerError, match="unsupported geometry"):
        BatchedLaneTracker(1, **bad)
    with pytest.raises(_lib.LaneTrackerError, match="n_average"):
        BatchedLaneTracker(1, **CAL, n_average=9)
    with pytest.raises(ValueError):
        BatchedLaneTracker(1, **dict(CAL, dist_coeffs=np.array([0.1] * 8)))
    free0 = torch_mod.cuda.mem_get_info()[0]
    S = 3
    b = BatchedLaneTracker(S, **CAL)
    vids = [synth.RoadVideo(40 + s) for s in range(S)]
    f0 = torch_mod.as_tensor(np.stack([v.frame(0) for v in vids])).cuda()
    with pytest.raises(ValueError):
        b.process(torch_mod.zeros((S + 1, 720, 1280, 3), dtype=torch_mod.uint8, device="cuda"))
    with pytest.raises(ValueError):
        b.process(f0.cpu())
    with pytest.raises(TypeError):
        b.process(f0, bogus_option=1)
    with pytest.raises(_lib.LaneTrackerError, match="window_width"):
        b.process(f0, window_width=65)
    # partial batch: only streams 0..1 advance
    n0 = lib.lt_launch_count()
    r = b.process(f0[:2].contiguous())
    assert len(r) == 2 and (r["counter"] == 1).all()
    assert 10 <= lib.lt_launch_count() - n0 <= 30
    assert b.get_state(2)[0].counter == 0 and b.get_state(1)[0].counter == 1
    b.process(f0)
    assert [b.get_state(s)[0].counter for s in range(S)] == [2, 2, 1]
    b.reset([1])
    assert [b.get_state(s)[0].counter for s in range(S)] == [2, 0, 1]
    assert b.get_state(1)[0].last_detection == 5
    with pytest.raises(_lib.LaneTrackerError):
        b.reset([7])
    # in-stream profiling adds up
    b.profile_begin(2)
    b.process(f0); b.process(f0)
    stages, calls = b.profile_read()
    assert calls == 2 and stages["erode55"] > 0 and stages["tophat55"] > 0 and stages["warp"] > 0
    # selective marks: only the named boundaries are recorded (and the second buffer set works on its own)
    b.profile_select(["warp", "erode55", "tophat55"])
    b.profile_begin(1)
    b.process_front_async(f0, 1)
    b.process_back_async(f0, None, 1)
    stages, calls = b.profile_read()
    assert calls == 1 and stages["erode55"] > 0 and stages["tophat55"] > 0 and stages["search"] == 0 and stages["cross_r"] == 0
    b.profile_select(None)
    with pytest.raises(_lib.LaneTrackerError, match="buffer set"):
        b.process_front_async(f0, 2)
    assert b.get_state(0)[0].counter == 5
    b.close()
    b.close()                                     # idempotent
    f2 = f0[:2].contiguous()
    torch_mod.cuda.synchronize()
    free1 = torch_mod.cuda.mem_get_info()[0]      # kernels loaded, torch's cache warm
    for _ in range(8):                            # handles release their device memory (both buffer sets)
        t = BatchedLaneTracker(2, **CAL)
        t.set_capture(True)
        t.process(f2)
        t.process_front_async(f2, 1)
        torch_mod.cuda.synchronize()
        t.close()
    torch_mod.cuda.synchronize()
    assert abs(torch_mod.cuda.mem_get_info()[0] - free1) < 16 << 20
    assert free0 > 0


def test_dropin_stage_methods_with_text(torch_mod, frames_np):
    """draw_lane / print_failure of the drop-in class include the reference's text overlays."""
    from lane_tracker_b200 import LaneTracker
    lt = LaneTracker(**CAL, print_frame_count=True)
    o = OracleLaneTracker(**CAL, print_frame_count=True, backend="cv2")
    vid = synth.RoadVideo(3)
    for t in range(2):
        a, b = lt.process(vid.frame(t)), o.process(vid.frame(t))
        assert _mism(a, b) == 0
    frame = frames_np[0]
    assert _mism(lt.draw_lane(frame), o.draw_lane(frame.copy())) == 0
    assert _mism(lt.print_failure(frame), o.print_failure(frame.copy())) == 0
    ly, lx, ry, rx = lt.get_poly_points(lt.left_avg_coeffs, lt.right_avg_coeffs, 1.0)
    assert np.array_equal(lx, o.left_avg_x) and np.array_equal(ry, o.right_avg_y)


def test_two_devices_in_one_process(torch_mod):
    """Handles on different GPUs of one process (per-device kernel attributes, no shared state)."""
    if torch_mod.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from lane_tracker_b200 import BatchedLaneTracker
    vid = synth.RoadVideo(50)
    frames = np.stack([vid.frame(t) for t in range(2)])
    outs = []
    for dev in (1, 0):
        with torch_mod.cuda.device(dev):
            b = BatchedLaneTracker(2, **CAL, device=dev)
            d = torch_mod.as_tensor(frames).to("cuda:%d" % dev)
            out = torch_mod.empty_like(d)
            res = b.process(d, out)
            outs.append((out.cpu().numpy(), res.copy()))
            b.close()
    assert np.array_equal(outs[0][0], outs[1][0]) and outs[0][1].tobytes() == outs[1][1].tobytes()


def _debug_view_frames(dv):
    vid = synth.RoadVideo(dv["seed"])
    outage = fx.load_frame(dv["outage_frame"])
    return [outage if kind == "outage" else vid.frame(t) for kind, t in dv["frames"]]


@pytest.mark.parametrize("mode", ["visualize_search", "split_view"])
def test_debug_views_match_reference_golden(torch_mod, mode):
    """process(visualize_search=True) / process(split_view=True) (lane_tracker.py:689-793, 1130-1209,
    utils.py:57-103) against the digests recorded from the reference: sliding-window view, band views, the
    attempt-2 view of an outage and the recovery, and the three-panel canvas with cv2.resize's fixed-point bilinear."""
    from lane_tracker_b200 import LaneTracker
    dv = GOLD["debug_views"]
    lt = LaneTracker(**CAL)
    for i, (frame, rec) in enumerate(zip(_debug_view_frames(dv), dv[mode])):
        before = frame.copy()
        r = lt.process(frame, **{mode: True})
        assert np.array_equal(frame, before)
        if mode == "visualize_search":
            assert fx.sha(r[0]) == rec["out"], i
            assert list(r[1].shape) == rec["vis_shape"] and fx.sha(r[1]) == rec["vis"], i
        else:
            assert list(r.shape) == rec["shape"] and fx.sha(r) == rec["canvas"], i


def test_debug_view_stage_methods_match_oracle(torch_mod, bt, frames_np):
    """The pieces behind the debug views against the CPU restatements, on inputs the golden sequence does not
    reach: bands that leave the canvas, windows clipped by the image border, up- and down-scaling resizes."""
    from lane_tracker_b200 import LaneTracker
    from lane_tracker_b200.utils import create_split_view
    rng = np.random.default_rng(21)
    # raw-frame bird's-eye view (lane_tracker.py:1035)
    got = bt.warp_frame(torch_mod.from_numpy(frames_np).cuda()).cpu().numpy()
    for i in (0, 3):
        assert np.array_equal(got[i], cvops.warp_perspective(frames_np[i], CAL["warp_matrices"][0], (1080, 1100)))
    # cv2.resize restatement: shrink, enlarge, identity, one channel
    for shape, dsize in (((1100, 1080, 3), (640, 652)), ((300, 500, 3), (777, 333)), ((64, 64), (100, 100)),
                         ((720, 1280, 3), (1280, 720)), ((301, 500, 3), (250, 150))):
        img = rng.integers(0, 256, shape, dtype=np.uint8)
        out = bt.resize_linear(torch_mod.from_numpy(img).cuda(), dsize).cpu().numpy()
        assert np.array_equal(out, cvops.resize_linear(img, dsize)), (shape, dsize)
    # search views on a random mask with hand-made search results
    lt = LaneTracker(**CAL)
    orc = OracleLaneTracker(**CAL)
    mask = (rng.random((1100, 1080)) < 0.05).astype(np.uint8) * 255
    for trial in range(6):
        n_l, n_r = int(rng.integers(1, 4000)), int(rng.integers(1, 4000))
        px = [rng.integers(0, 1100, n_l), rng.integers(0, 1080, n_l), rng.integers(0, 1100, n_r), rng.integers(0, 1080, n_r)]
        cents = [[int(v) for v in rng.integers(-20, 1100, int(rng.integers(0, 30)))] for _ in range(2)]
        for o in (lt, orc):
            o.left_y, o.left_x, o.right_y, o.right_x = px
            o.left_window_centroids, o.right_window_centroids = cents
            rng2 = np.random.default_rng(100 + trial)          # same coefficients for both objects
            o.last_left_coeffs = np.array([rng2.normal(0, 2e-4), rng2.normal(0, 0.3), rng2.uniform(-20, 300)])
            o.last_right_coeffs = np.array([rng2.normal(0, 2e-4), rng2.normal(0, 0.3), rng2.uniform(800, 1100)])
        lf = np.array([rng.normal(0, 1e-4), rng.normal(0, 0.2), rng.uniform(200, 500)])
        rf = lf + np.array([0.0, 0.0, 180.0])
        ww, wh, ib = int(rng.choice([30, 60, 200])), int(rng.choice([40, 25])), int(rng.choice([30, 0]))
        assert np.array_equal(lt.visualize_sliding_window_search(mask, lf, rf, ww, wh, ib),
                              orc.visualize_sliding_window_search(mask, lf, rf, ww, wh, ib)), trial
        bw, partial = int(rng.choice([25, 60, 150])), float(rng.choice([1.0, 0.5, 0.3]))
        assert np.array_equal(lt.visualize_band_search(mask, lf, rf, bw, partial),
                              orc.visualize_band_search(mask, lf, rf, bw, partial)), trial
    # split view helpers, NumPy in / NumPy out
    a, b, c = frames_np[0], rng.integers(0, 256, (1100, 1080, 3), dtype=np.uint8), rng.integers(0, 256, (1100, 1080, 3), dtype=np.uint8)
    assert np.array_equal(lt.triple_split_view([a, b, c]), orc.triple_split_view([a, b, c]))
    assert np.array_equal(create_split_view((900, 500), [b, a], [(10, 20), (300, 100)], [(400, 300), (640, 360)]),
                          orc.create_split_view((900, 500), [b, a], [(10, 20), (300, 100)], [(400, 300), (640, 360)]))
    with pytest.raises(ValueError):                            # a 2-D third panel cannot be placed: reference behaviour
        lt.triple_split_view([a, b, mask])
    with pytest.raises(NotImplementedError):
        create_split_view((900, 500), [b], [(0, 0)], [(400, 300)], captions=["x"])


def test_device_pipeline_matches_sequential_process(torch_mod):
    """DevicePipeline (front half of batch k+1 on one stream under the back half of batch k on another, two
    intermediate buffer sets) == sequential process() calls: results, state and annotated frames, including a
    stream that fails every attempt (second-attempt filter reads the planes of its own buffer set)."""
    from lane_tracker_b200 import BatchedLaneTracker, DevicePipeline
    S, T = 3, 7
    vids = [synth.RoadVideo(s) for s in range(S - 1)]
    noise = np.random.default_rng(5).integers(0, 256, (720, 1280, 3), dtype=np.uint8)
    batches = [torch_mod.from_numpy(np.stack([v.frame(t) for v in vids] + [noise])).cuda() for t in range(T)]
    seq = BatchedLaneTracker(S, **CAL)
    want = []
    for b in batches:
        out = torch_mod.empty_like(b)
        want.append((seq.process(b, out), out))
    pip = BatchedLaneTracker(S, **CAL)
    dp = DevicePipeline(pip)
    outs = [torch_mod.empty_like(b) for b in batches]
    got = []
    for b, o in zip(batches, outs):
        dp.submit(b, o)
        got.append(dp.fetch_results(S))
    # and without synchronising between submissions
    pip2 = BatchedLaneTracker(S, **CAL)
    dp2 = DevicePipeline(pip2)
    outs2 = [torch_mod.empty_like(b) for b in batches]
    for b, o in zip(batches, outs2):
        dp2.submit(b, o)
    last = dp2.fetch_results(S)
    for t in range(T):
        for f in want[t][0].dtype.names:
            assert np.array_equal(want[t][0][f], got[t][f]), (t, f)
        assert torch_mod.equal(want[t][1], outs[t]) and torch_mod.equal(want[t][1], outs2[t]), t
    for f in last.dtype.names:
        assert np.array_equal(want[-1][0][f], last[f]), f
    assert want[-1][0]["attempts"][S - 1] == 2
    for s in range(S):
        a, b = seq.get_state(s), pip2.get_state(s)
        assert bytes(a[0]) == bytes(b[0]) and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
    for t_ in (seq, pip, pip2):
        t_.close()


def test_graphed_process_matches_sequential_process(torch_mod):
    """GraphedProcess (lt_process captured once into a CUDA graph, side-stream fork/join included) == process():
    results, frames and state; building the graph leaves the tracking state untouched."""
    from lane_tracker_b200 import BatchedLaneTracker, GraphedProcess
    S, T = 2, 6
    vids = [synth.RoadVideo(20 + s) for s in range(S)]
    batches = [torch_mod.from_numpy(np.stack([v.frame(t) for v in vids])).cuda() for t in range(T)]
    seq, gr = BatchedLaneTracker(S, **CAL), BatchedLaneTracker(S, **CAL)
    want = []
    for b in batches:
        o = torch_mod.empty_like(b)
        want.append((seq.process(b, o), o))
    gr.process(batches[0], torch_mod.empty_like(batches[0]))        # some state before the capture
    g = GraphedProcess(gr, S)
    assert gr.get_state(0)[0].counter == 1 and bytes(gr.get_state(1)[0]) == bytes(seq_state_after_one(seq, CAL, batches[0], 1))
    for t in range(1, T):
        g.frames.copy_(batches[t], non_blocking=True)
        g.replay()
        r = g.fetch_results()
        for f in r.dtype.names:
            assert np.array_equal(r[f], want[t][0][f]), (t, f)
        assert torch_mod.equal(g.out, want[t][1]), t
    for s in range(S):
        a, b = seq.get_state(s), gr.get_state(s)
        assert bytes(a[0]) == bytes(b[0]) and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
    seq.close(); gr.close()


def seq_state_after_one(_seq, cal, batch, stream):
    """lt_state of `stream` after exactly one process() call on a fresh tracker."""
    from lane_tracker_b200 import BatchedLaneTracker
    t = BatchedLaneTracker(int(batch.shape[0]), **cal)
    t.process(batch)
    st = t.get_state(stream)[0]
    t.close()
    return st
