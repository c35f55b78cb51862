"""GPU parity at the configurations the benchmark runs (BASELINE.json configs[1]-[3]) and for the code paths the
first-round suite left uncovered: the 64-stream band geometry of the morphology launch, wide / unpacked cross
threshold kernels, mask_noise with small kernels, find_lane_points() with its own defaults, large band widths.
Everything goes through the C ABI; the CPU oracle is the checker."""
import os
import warnings

import numpy as np
import pytest

import _fixtures as fx
import _oracle_pool
from lane_tracker_b200 import synth
from oracle.tracker import OracleLaneTracker

pytestmark = pytest.mark.gpu

CAL = synth.shipped_calibration()
GOLD = fx.golden()
FIT_RTOL = 1e-6


@pytest.fixture(scope="module")
def torch_mod():
    import torch
    assert torch.cuda.is_available()
    return torch


def _mism(a, b):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, (a.shape, b.shape)
    return int((a != b).sum())


def _check_stream_frame(res, want, tag):
    assert int(res["counter"]) == want["counter"], tag
    assert int(res["success"]) == want["success"], tag
    assert int(res["attempts"]) == want["attempts"], tag
    assert int(res["search_mode"]) == (1 if want["mode"] == "bs" else 0), tag
    assert bool(res["detected_pixels"]) == want["detected"], tag
    assert bool(res["valid_lane_lines"]) == want["valid"], tag
    assert int(res["last_detection"]) == want["last_detection"], tag
    assert bool(res["first_detected"]) == want["first_detected"] and bool(res["first_valid"]) == want["first_valid"], tag
    if want["detected"]:
        assert int(res["n_left"]) == want["n_left"] and int(res["n_right"]) == want["n_right"], tag
        np.testing.assert_allclose(res["left_fit"], want["left_fit"], rtol=FIT_RTOL, err_msg=str(tag))
        np.testing.assert_allclose(res["right_fit"], want["right_fit"], rtol=FIT_RTOL, err_msg=str(tag))
    if want["valid"]:
        assert int(res["average_curve_radius"]) == want["radius"], tag
        assert float(res["eccentricity"]) == pytest.approx(want["ecc"], rel=1e-12, abs=1e-15), tag
        np.testing.assert_allclose(res["left_avg"], want["left_avg"], rtol=FIT_RTOL, err_msg=str(tag))


def test_benchmark_configuration_64_streams(torch_mod):
    """BASELINE.json configs[2] as bench.py runs it: 64 streams per batch through DevicePipeline -- 53 synthetic road
    videos plus the 11 bundled photographs (which take both attempts and the sliding-window search every frame) --
    six frames each.  Every stream's result record, every output frame and the final masks against the oracle; the
    bundled frames also against the digests recorded from the reference itself.  Asserts that the morphology ran
    with the band split of a 64-stream launch (2 bands of 550 rows for the 55x55 kernel on 148 SMs)."""
    from lane_tracker_b200 import BatchedLaneTracker, DevicePipeline
    torch = torch_mod
    names = fx.frame_names()
    n_syn, T = 64 - len(names), 6
    jobs = [dict(kind="synth", seed=s, frames=T) for s in range(n_syn)] + [dict(kind="image", name=n, frames=T) for n in names]
    want = _oracle_pool.run_jobs(jobs)
    photos = np.stack([fx.load_frame(n) for n in names])
    vids = [synth.RoadVideo(s) for s in range(n_syn)]
    os.environ.pop("LT_MORPH_BANDS", None)
    bt = BatchedLaneTracker(64, **CAL)
    pipe = DevicePipeline(bt)
    outs = [torch.empty((64, 720, 1280, 3), dtype=torch.uint8, device="cuda") for _ in range(2)]
    for t in range(T):
        batch = np.concatenate([np.stack([v.frame(t) for v in vids]), photos])
        d = torch.as_tensor(batch).cuda()
        pipe.submit(d, outs[t & 1])
        res = pipe.fetch_results(64)
        out = outs[t & 1].cpu().numpy()
        for s in range(64):
            tag = (t, s)
            _check_stream_frame(res[s], want[s][t], tag)
            assert fx.sha(out[s]) == want[s][t]["out"], tag
        if t == 0:
            for i, n in enumerate(names):
                assert fx.sha(out[n_syn + i]) == GOLD["images"][n]["process_out_full"], n
    for s in range(64):
        assert fx.sha(bt.debug_read("mask", s)) == want[s][T - 1]["mask"], s
    for i, n in enumerate(names):      # second-attempt mask of a bundled frame, as recorded from the reference
        assert fx.sha(bt.debug_read("mask", n_syn + i)) == GOLD["images"][n]["mask_neighborhood"], n
    bands = bt.morph_bands()
    if bt.sm_count == 148:          # the band split of the benchmarked launch: 2 bands of 550 rows for the 55x55 kernel
        assert bands[0] == 2 and bands[1] in (5, 6), bands
    bt.close()


@pytest.mark.parametrize("bands", ["1,1", "2,5", "3,7", "24,10"])
def test_tophat_band_geometries(torch_mod, bands):
    """The ellipse top-hats under every band split the launcher can choose (LT_MORPH_BANDS pins it): bit-exact."""
    from lane_tracker_b200 import BatchedLaneTracker
    torch = torch_mod
    rng = np.random.default_rng(21)
    frames = np.stack([fx.load_frame("test4.jpg"), synth.RoadVideo(5).frame(2),
                       rng.integers(0, 256, (720, 1280, 3), dtype=np.uint8)])
    os.environ["LT_MORPH_BANDS"] = bands
    try:
        bt = BatchedLaneTracker(3, **CAL)
        bt.remap(torch.as_tensor(frames).cuda(), want_bv=False)
        bt.filter_lane_points(None, "bilateral", 15, 8, 35, 5)
        assert bt.morph_bands() == tuple(int(v) for v in bands.split(","))
        for i, f in enumerate(frames):
            o = OracleLaneTracker(**CAL)
            bv = o._remap(f)
            want = o.filter_lane_points(bv, "bilateral", 15, 8, 35, 5, False, 65, 10, 140)
            assert _mism(bt.debug_read("r_tophat", i), o.trace["r_tophat"]) == 0, (bands, i)
            assert _mism(bt.debug_read("b_tophat", i), o.trace["b_tophat"]) == 0, (bands, i)
            assert _mism(bt.debug_read("mask", i), want) == 0, (bands, i)
        bt.close()
    finally:
        os.environ.pop("LT_MORPH_BANDS", None)


@pytest.mark.parametrize("k,C", [(130, 2), (200, 1), (255, 0), (100, 80), (127, 3), (90, 110)])
def test_cross_threshold_wide_and_unpacked_kernels(torch_mod, k, C):
    """Window sizes beyond the packed 15-bit lanes: k > 127 (prefix-sum row kernel) and k*255 + C*k >= 2^15 (unpacked
    column kernel), including windows longer than the row padding of the planes."""
    from lane_tracker_b200 import BatchedLaneTracker
    torch = torch_mod
    bv = OracleLaneTracker(**CAL)._remap(fx.load_frame("test2.jpg"))
    o = OracleLaneTracker(**CAL)
    want = o.filter_lane_points(bv, "bilateral", k, C, max(1, k - 9), C + 1, True, k, C, 120)
    bt = BatchedLaneTracker(1, **CAL)
    got = bt.filter_lane_points(torch.as_tensor(bv[None]).cuda(), "bilateral", k, C, max(1, k - 9), C + 1, True, k, C, 120)
    assert _mism(got.cpu().numpy()[0], want) == 0
    assert _mism(bt.debug_read("merged", 0), o.trace["merged"]) == 0
    bt.close()


@pytest.mark.parametrize("kn,Cn,thr", [(20, 10, 140), (40, 10, 140), (33, 4, 128), (65, 10, 135), (48, 0, 150)])
def test_mask_noise_kernel_sizes(torch_mod, kn, Cn, thr):
    """mask_noise (lane_tracker.py:221-231) thresholds the RAW Lab-b plane, whose pad rows hold the erosion pad and
    not the filter's zero border: every window size must take the bounds-checked column walk."""
    from lane_tracker_b200 import BatchedLaneTracker
    torch = torch_mod
    rng = np.random.default_rng(4)
    bvs = [OracleLaneTracker(**CAL)._remap(fx.load_frame("test5.jpg")),
           rng.integers(0, 256, (1100, 1080, 3), dtype=np.uint8)]
    bvs[1][:60] = 255          # bright top / bottom rows: the case the pad rows would have corrupted
    bvs[1][-60:] = (250, 250, 20)
    bt = BatchedLaneTracker(2, **CAL)
    got = bt.filter_lane_points(torch.as_tensor(np.stack(bvs)).cuda(), "bilateral", 15, 8, 35, 5, True, kn, Cn, thr).cpu().numpy()
    for i, bv in enumerate(bvs):
        want = OracleLaneTracker(**CAL).filter_lane_points(bv, "bilateral", 15, 8, 35, 5, True, kn, Cn, thr)
        assert _mism(got[i], want) == 0, i
    bt.close()


def test_find_lane_points_method_with_its_own_defaults(torch_mod):
    """LaneTracker.find_lane_points (lane_tracker.py:795-874): defaults mask_noise=True, bandwidth=30, partial=0.5."""
    from lane_tracker_b200 import LaneTracker
    warnings.simplefilter("ignore")
    lt = LaneTracker(**CAL)
    o = OracleLaneTracker(**CAL)
    vid = synth.RoadVideo(7)
    for f in (fx.load_frame("test3.jpg"), vid.frame(0)):
        mask, mode = lt.find_lane_points(f)
        wmask, wmode = o.find_lane_points(f)
        assert mode == wmode == "sws" and _mism(mask, wmask) == 0
        assert lt.detected_pixels == o.detected_pixels
        if o.detected_pixels:
            assert np.array_equal(lt.left_x, o.left_x) and np.array_equal(lt.left_y, o.left_y)
            assert np.array_equal(lt.right_x, o.right_x) and np.array_equal(lt.right_y, o.right_y)
            assert lt.left_window_centroids == o.left_window_centroids
            assert lt.right_window_centroids == o.right_window_centroids
    # after a valid process() call the method switches to the band search around the last fit
    f0, f1 = vid.frame(0), vid.frame(1)
    lt.process(f0)
    o.process(f0.copy())
    assert lt.valid_lane_lines and o.valid_lane_lines
    mask, mode = lt.find_lane_points(f1)
    wmask, wmode = o.find_lane_points(f1)
    assert mode == wmode == "bs" and _mism(mask, wmask) == 0
    assert np.array_equal(lt.left_x, o.left_x) and np.array_equal(lt.right_y, o.right_y)
    # explicit options
    mask, mode = lt.find_lane_points(f1, filter_type="neighborhood", mask_noise=False, bandwidth=45, partial=1.0)
    wmask, wmode = o.find_lane_points(f1, filter_type="neighborhood", mask_noise=False, bandwidth=45, partial=1.0)
    assert mode == wmode and _mism(mask, wmask) == 0
    assert np.array_equal(lt.left_x, o.left_x) and np.array_equal(lt.right_x, o.right_x)
    with pytest.raises(ValueError):
        lt.find_lane_points(f1, filter_type="gaussian")


@pytest.mark.parametrize("bandwidth", [33, 40, 64])
def test_band_search_wider_than_the_default_pixel_capacity(torch_mod, bandwidth):
    """A band-search row holds up to 2*bandwidth - 1 pixels: beyond bandwidth 32 the default capture capacity
    (64 per row) would truncate the pixel lists; the drop-in raises the capacity, the C ABI refuses."""
    from lane_tracker_b200 import BatchedLaneTracker, LaneTracker, _lib
    warnings.simplefilter("ignore")
    torch = torch_mod
    lt = LaneTracker(**CAL)
    o = OracleLaneTracker(**CAL)
    vid = synth.RoadVideo(3)
    for t in range(3):
        f = vid.frame(t)
        out = lt.process(f, bandwidth=bandwidth)
        want = o.process(f.copy(), bandwidth=bandwidth)
        assert _mism(out, want) == 0, t
        assert np.array_equal(lt.left_x, o.left_x) and np.array_equal(lt.left_y, o.left_y), t
        assert np.array_equal(lt.right_x, o.right_x) and np.array_equal(lt.right_y, o.right_y), t
    assert lt.last_result["search_mode"] == 1
    bt = BatchedLaneTracker(1, **CAL)
    bt.set_capture(True)
    d = torch.as_tensor(vid.frame(0)[None]).cuda()
    with pytest.raises(_lib.LaneTrackerError):
        bt.process(d, None, bandwidth=bandwidth)
    bt.set_pixel_capacity((2 * bandwidth - 1) * 1100)
    bt.process(d, None, bandwidth=bandwidth)
    bt.close()


def test_draw_lane_stage_call_keeps_the_cached_lane_polygon(torch_mod):
    """lt_draw_lane must not overwrite the polygon process() caches for the frames that fail within n_fail
    (lane_tracker.py:1160-1166)."""
    from lane_tracker_b200 import LaneTracker
    warnings.simplefilter("ignore")
    lt = LaneTracker(**CAL)
    o = OracleLaneTracker(**CAL)
    vid = synth.RoadVideo(9)
    for t in range(2):
        lt.process(vid.frame(t))
        o.process(vid.frame(t))
    # a stage call with a very different polygon, on both sides
    saved = (lt.left_avg_x.copy(), lt.right_avg_x.copy())
    lt.left_avg_x = np.full_like(saved[0], 100)
    lt.right_avg_x = np.full_like(saved[1], 900)
    lt.draw_lane(vid.frame(1))
    lt.left_avg_x, lt.right_avg_x = saved
    blank = np.full((720, 1280, 3), 90, np.uint8)       # no lane pixels: the cached polygon is drawn again
    got = lt.process(blank)
    want = o.process(blank.copy())
    assert not lt.valid_lane_lines and not o.valid_lane_lines
    assert _mism(got, want) == 0


def test_parameter_sets_the_kernels_cannot_honour_are_rejected(torch_mod):
    from lane_tracker_b200 import BatchedLaneTracker, _lib
    torch = torch_mod
    bt = BatchedLaneTracker(1, **CAL)
    d = torch.zeros((1, 720, 1280, 3), dtype=torch.uint8, device="cuda")
    with pytest.raises(_lib.LaneTrackerError, match="search levels"):
        bt.process(d, None, window_height=5)            # 214 levels > LT_MAX_LEVELS
    bt.process(d, None, window_height=9)                # 118 levels
    for bad in (dict(search_range=-1), dict(ignore_sides=5000), dict(bandwidth=-3), dict(start_slice=1.5)):
        with pytest.raises(_lib.LaneTrackerError):
            bt.process(d, None, **bad)
    bt.close()


def test_dropin_metric_and_text_methods_on_the_device(torch_mod):
    """get_curve_radius / get_eccentricity (lane_tracker.py:530-559), draw_lane / print_failure with their putText overlays
    (629-673) and the module-level bilateral_adaptive_threshold (14-83) as drop-in calls, against the oracle / cv2."""
    import lane_tracker_b200 as ltb
    from lane_tracker_b200 import LaneTracker
    warnings.simplefilter("ignore")
    lt = LaneTracker(**CAL, print_frame_count=True)
    o = OracleLaneTracker(**CAL, print_frame_count=True)
    vid = synth.RoadVideo(12)
    for t in range(3):
        f = vid.frame(t)
        assert _mism(lt.process(f), o.process(f.copy())) == 0
    # the two metric methods, called directly on the state process() left behind
    lt.average_curve_radii, o.average_curve_radii = list(lt.average_curve_radii), list(o.average_curve_radii)
    lt.get_curve_radius()
    o.get_curve_radius()
    assert (lt.left_curve_radius, lt.right_curve_radius) == (o.left_curve_radius, o.right_curve_radius)
    assert lt.average_curve_radii == o.average_curve_radii and lt.average_curve_radius == o.average_curve_radius
    lt.eccentricity = None
    lt.get_eccentricity()
    o.get_eccentricity()
    assert lt.eccentricity == o.eccentricity
    # draw_lane / print_failure: text + polygon, the caller's frame untouched
    f = vid.frame(5)
    keep = f.copy()
    assert _mism(lt.draw_lane(f), o.draw_lane(f.copy())) == 0
    assert _mism(lt.print_failure(f), o.print_failure(f.copy())) == 0
    assert np.array_equal(f, keep)
    # module-level threshold, both modes, other mask values, a non-bird's-eye image size
    import cv2
    rng = np.random.default_rng(8)
    img = cv2.GaussianBlur(rng.integers(0, 256, (333, 517), dtype=np.uint8), (0, 0), 2.0)
    for kw in (dict(), dict(ksize=12, C=3), dict(ksize=45, C=2, mode='ceil'), dict(ksize=7, C=0, true_value=200, false_value=9)):
        want = _reference_bilateral(img, **kw)
        assert _mism(ltb.bilateral_adaptive_threshold(img, **kw), want) == 0, kw
    with pytest.raises(ValueError):
        ltb.bilateral_adaptive_threshold(img, mode='round')


def _reference_bilateral(img, ksize=30, C=0, mode='floor', true_value=255, false_value=0):
    """The reference's own filter2D formulation (lane_tracker.py:61-81), executed with cv2."""
    import cv2
    mask = np.full(img.shape, false_value, dtype=np.uint8)
    kl = np.array([[1] * ksize + [-ksize]], dtype=np.int16)
    kr = np.array([[-ksize] + [1] * ksize], dtype=np.int16)
    ku = np.array([[1]] * ksize + [[-ksize]], dtype=np.int16)
    kd = np.array([[-ksize]] + [[1]] * ksize, dtype=np.int16)
    delta = C * ksize if mode == 'floor' else -C * ksize
    lt_ = cv2.filter2D(img, cv2.CV_16S, kl, anchor=(ksize, 0), delta=delta, borderType=cv2.BORDER_CONSTANT)
    rt_ = cv2.filter2D(img, cv2.CV_16S, kr, anchor=(0, 0), delta=delta, borderType=cv2.BORDER_CONSTANT)
    ut_ = cv2.filter2D(img, cv2.CV_16S, ku, anchor=(0, ksize), delta=delta, borderType=cv2.BORDER_CONSTANT)
    dt_ = cv2.filter2D(img, cv2.CV_16S, kd, anchor=(0, 0), delta=delta, borderType=cv2.BORDER_CONSTANT)
    if mode == 'floor':
        mask[((0 > lt_) & (0 > rt_)) | ((0 > ut_) & (0 > dt_))] = true_value
    else:
        mask[((0 < lt_) & (0 < rt_)) | ((0 < ut_) & (0 < dt_))] = true_value
    return mask


def test_nv12_ingest_matches_oracle_and_feeds_process(torch_mod):
    """lt_nv12_to_rgb: decoder output -> RGB frames, bit-exact with the oracle restatement of
    cv2.cvtColor(COLOR_YUV2RGB_NV12); the converted frames then go through process() like any other frame."""
    from lane_tracker_b200 import BatchedLaneTracker
    from oracle import cvops
    torch = torch_mod
    rng = np.random.default_rng(5)
    H, W = 720, 1280
    nv = rng.integers(0, 256, (3, H * 3 // 2, W), dtype=np.uint8)
    nv[0, :2, :8] = [[0, 255, 16, 235, 15, 17, 1, 254]] * 2          # luma extremes
    nv[0, H, :8] = [0, 0, 255, 255, 0, 255, 255, 0]                  # chroma extremes: every channel saturates
    # frame 2: a rendered road frame taken through an RGB -> NV12 round trip (what a camera / decoder delivers)
    vid = synth.RoadVideo(3)
    rgb = vid.frame(0).astype(np.float64)
    y = 16 + (65.481 * rgb[..., 0] + 128.553 * rgb[..., 1] + 24.966 * rgb[..., 2]) / 255
    cb = 128 + (-37.797 * rgb[..., 0] - 74.203 * rgb[..., 1] + 112.0 * rgb[..., 2]) / 255
    cr = 128 + (112.0 * rgb[..., 0] - 93.786 * rgb[..., 1] - 18.214 * rgb[..., 2]) / 255
    nv[2, :H] = np.clip(np.rint(y), 0, 255)
    sub = lambda c: c.reshape(H // 2, 2, W // 2, 2).mean(axis=(1, 3))
    nv[2, H:] = np.clip(np.rint(np.stack([sub(cb), sub(cr)], axis=-1)), 0, 255).reshape(H // 2, W)
    trk = BatchedLaneTracker(3, **CAL, device=0)
    try:
        got = trk.nv12_to_rgb(torch.as_tensor(nv).cuda())
        want = np.stack([cvops.yuv2rgb_nv12(f, W, H) for f in nv])
        assert _mism(got.cpu().numpy(), want) == 0
        with pytest.raises(ValueError):
            trk.nv12_to_rgb(torch.zeros((1, 1081, 1280), dtype=torch.uint8, device="cuda"))
        # the converted frames are ordinary input frames: same result records as the oracle on the same RGB
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            res = trk.process(got, n_tries=1)
            ref = OracleLaneTracker(**CAL)
            ref.process(want[2].copy(), n_tries=1)
        assert bool(res["detected_pixels"][2]) == bool(ref.detected_pixels)
        assert bool(res["valid_lane_lines"][2]) == bool(ref.valid_lane_lines)
        assert _mism(trk.debug_read("mask", 2), ref.trace["attempts"][0]["mask"]) == 0
    finally:
        trk.close()


def test_sliding_window_search_sees_the_leftmost_columns(torch_mod):
    """Regression (round-2 racecheck): the histogram of the last mask word of a level used to spill into the
    prefix-sum row of the next level and could zero the counts of columns 0..6.  With ignore_sides = 0 and a lane line
    that hugs the left image edge those columns decide the windowed counts (and, closer to the edge, NumPy's negative
    slice start empties the window: the oracle decides, the device must agree either way)."""
    from lane_tracker_b200 import BatchedLaneTracker
    torch = torch_mod
    kw = dict(window_width=30, window_height=40, search_range=20, mu=0.1, no_success_limit=8, start_slice=0.25,
              ignore_sides=0, ignore_bottom=30, partial=1.0)
    trk = BatchedLaneTracker(1, **CAL, device=0)
    detected = 0
    try:
        for x0, w in ((0, 5), (0, 12), (2, 20), (4, 24), (0, 30), (6, 20), (10, 16)):
            mask = np.zeros((1100, 1080), np.uint8)
            mask[:, x0:x0 + w] = 255            # left line next to the image edge
            mask[:, 1050:1080] = 255            # right line in the last (partial) mask word
            o = OracleLaneTracker(**CAL)
            o.sliding_window_search(mask, **kw)
            for _ in range(2):                  # the race was timing dependent
                px, cents, det = trk.sliding_window_search(torch.as_tensor(mask[None]).cuda(), 30, 40, 20, 0.1, 8, 0.25, 0, 30, 1.0)
                assert bool(det[0]) == o.detected_pixels, (x0, w)
                assert cents[0][0] == o.trace["sws_centroids"][0] and cents[0][1] == o.trace["sws_centroids"][1], (x0, w)
                if o.detected_pixels:
                    (ly, lx), (ry, rx) = px[0]
                    assert np.array_equal(ly, o.left_y) and np.array_equal(lx, o.left_x), (x0, w)
                    assert np.array_equal(ry, o.right_y) and np.array_equal(rx, o.right_x), (x0, w)
            detected += bool(o.detected_pixels)
        assert detected >= 3
    finally:
        trk.close()


def test_stream_counts_that_straddle_the_16_stream_groups(torch_mod):
    """The undistorted buffer is organised in groups of 16 streams and chunks of 4 (lt_remap.cu): stream counts that end
    inside a chunk of the second group must give every stream the same result as the same frame in slot 0 / 1."""
    from lane_tracker_b200 import BatchedLaneTracker
    torch = torch_mod
    vid = synth.RoadVideo(7)
    a, b = vid.frame(0), vid.frame(1)
    for S in (18, 21):
        frames = np.stack([a if s % 2 == 0 else b for s in range(S)])
        trk = BatchedLaneTracker(S, **CAL, device=0)
        try:
            bv = trk.remap(torch.as_tensor(frames).cuda()).cpu().numpy()
            for s in range(2, S):
                assert _mism(bv[s], bv[s % 2]) == 0, (S, s)
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                res = trk.process(torch.as_tensor(frames).cuda(), n_tries=1)
            for s in range(2, S):
                assert _mism(trk.debug_read("mask", s), trk.debug_read("mask", s % 2)) == 0, (S, s)
                assert res["n_left"][s] == res["n_left"][s % 2] and res["n_right"][s] == res["n_right"][s % 2], (S, s)
            # and slot 0 itself against the oracle
            ref = OracleLaneTracker(**CAL)
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                ref.process(a.copy(), n_tries=1)
            assert _mism(trk.debug_read("mask", 0), ref.trace["attempts"][0]["mask"]) == 0
        finally:
            trk.close()
