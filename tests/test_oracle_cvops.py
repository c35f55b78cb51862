"""Pin oracle/cvops.py bit-for-bit against cv2 itself (the reference's un-vendored kernel library)."""
import numpy as np
import pytest

from oracle import cvops
from lane_tracker_b200 import synth

cv2 = pytest.importorskip("cv2")

CAL = synth.shipped_calibration()
K, D = CAL["cam_matrix"], CAL["dist_coeffs"]
M, MINV = CAL["warp_matrices"]


@pytest.fixture(scope="module")
def noise():
    return np.random.default_rng(7).integers(0, 256, (720, 1280, 3), dtype=np.uint8)


@pytest.fixture(scope="module")
def bv(noise):
    und = cv2.undistort(noise, K, D, None, K)
    return cv2.warpPerspective(und, M, (1080, 1100), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT)


def test_undistort(noise):
    assert np.array_equal(cvops.undistort(noise, K, D), cv2.undistort(noise, K, D, None, K))


def test_warp_and_unwarp(noise, bv):
    und = cv2.undistort(noise, K, D, None, K)
    assert np.array_equal(cvops.warp_perspective(und, M, (1080, 1100)), bv)
    assert np.array_equal(cvops.warp_perspective(bv, MINV, (1280, 720)),
                          cv2.warpPerspective(bv, MINV, (1280, 720)))


def test_unwarp_tie_pixel_needs_cv_invert():
    """Frame pixel (707, 858) maps to X = 18867.5 exactly with cv::invert's adjugate inverse of Minv
    (rounds to even, 18868) but to 18867.49999999999 with LAPACK's; its row lies below the 1100-row
    canvas, so only a taller source exposes which one OpenCV uses."""
    Xc, Yc = cvops.perspective_map_q5(MINV, 1280, 720)
    assert (Xc[707, 858], Yc[707, 858] >> 5) == (18868, 1107)
    src = np.random.default_rng(0).integers(0, 256, (1300, 1080), dtype=np.uint8)
    want = cv2.warpPerspective(src, MINV, (1280, 720), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT)
    assert np.array_equal(cvops.bilinear_q5(src, Xc, Yc), want)
    src = np.random.default_rng(1).integers(0, 256, (900, 1500), dtype=np.uint8)
    Xm, Ym = cvops.perspective_map_q5(M, 1080, 1100)
    want = cv2.warpPerspective(src, M, (1080, 1100), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT)
    assert np.array_equal(cvops.bilinear_q5(src, Xm, Ym), want)


def test_lab_b_exhaustive_slice():
    # every (r,g) pair for 16 blue levels = 1M colours; the full 2^24 was checked in the survey
    r, g, b = np.meshgrid(np.arange(256), np.arange(256), np.arange(0, 256, 17)[:16], indexing="ij")
    rgb = np.stack([r, g, b], axis=-1).astype(np.uint8).reshape(256, -1, 3)
    assert np.array_equal(cvops.lab_b_plane(rgb), cv2.cvtColor(rgb, cv2.COLOR_RGB2LAB)[:, :, 2])


@pytest.mark.parametrize("k", [5, 29, 55])
def test_ellipse_morphology(bv, k):
    se = cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (k, k))
    assert [int(v) for v in (se.sum(1) - 1) // 2] == cvops.ellipse_half_widths(k)
    smooth = cv2.GaussianBlur(bv[:, :, 1], (0, 0), 3)
    for plane in (bv[:, :, 0], smooth):
        assert np.array_equal(cvops.erode_ellipse(plane, k), cv2.erode(plane, se))
        assert np.array_equal(cvops.dilate_ellipse(plane, k), cv2.dilate(plane, se))
        assert np.array_equal(cvops.tophat_ellipse(plane, k), cv2.morphologyEx(plane, cv2.MORPH_TOPHAT, se))
        assert np.array_equal(cvops.open_ellipse(plane, k), cv2.morphologyEx(plane, cv2.MORPH_OPEN, se))


@pytest.mark.parametrize("k,C", [(15, 8), (35, 5), (65, 10), (25, 8)])
def test_cross_threshold(bv, k, C):
    img = cv2.GaussianBlur(bv[:, :, 2], (0, 0), 2)
    kl = np.array([[1] * k + [-k]], dtype=np.int16)
    kr = np.array([[-k] + [1] * k], dtype=np.int16)
    d = C * k
    a = cv2.filter2D(img, cv2.CV_16S, kl, anchor=(k, 0), delta=d, borderType=cv2.BORDER_CONSTANT)
    b = cv2.filter2D(img, cv2.CV_16S, kr, anchor=(0, 0), delta=d, borderType=cv2.BORDER_CONSTANT)
    c = cv2.filter2D(img, cv2.CV_16S, kl.T.copy(), anchor=(0, k), delta=d, borderType=cv2.BORDER_CONSTANT)
    e = cv2.filter2D(img, cv2.CV_16S, kr.T.copy(), anchor=(0, 0), delta=d, borderType=cv2.BORDER_CONSTANT)
    want = np.where(((a < 0) & (b < 0)) | ((c < 0) & (e < 0)), 255, 0).astype(np.uint8)
    got = cvops.cross_threshold(img, k, C)
    assert 0 < (got > 0).mean() < 1
    assert np.array_equal(got, want)


@pytest.mark.parametrize("bs,c", [(15, 5), (35, 5), (15, 8)])
def test_box_mean_threshold(bv, bs, c):
    for plane in (bv[:, :, 0], cv2.GaussianBlur(bv[:, :, 1], (0, 0), 2)):
        want = cv2.adaptiveThreshold(plane, 255, cv2.ADAPTIVE_THRESH_MEAN_C, cv2.THRESH_BINARY, bs, -c)
        assert np.array_equal(cvops.box_mean_threshold(plane, bs, c), want)


def test_add_weighted(noise):
    lane = np.random.default_rng(8).integers(0, 256, noise.shape, dtype=np.uint8)
    assert np.array_equal(cvops.add_weighted_03(noise, lane), cv2.addWeighted(noise, 1, lane, 0.3, 0))


def test_fill_poly_lane_polygons():
    W, H = 1080, 1100
    rng = np.random.default_rng(1)
    checked = 0
    for t in range(120):
        a = rng.normal(0, 3e-4) * (5 if t % 3 == 0 else 1)
        b = rng.normal(0, 0.5) * (3 if t % 3 == 0 else 1)
        c = rng.uniform(100, 700)
        a2, b2, sep = a + rng.normal(0, 1e-4), b + rng.normal(0, 0.1), rng.uniform(100, 300)
        partial = 1.0 if t % 2 == 0 else 0.5
        ploty = np.linspace(H * (1 - partial), H - 1, int(H * partial))
        xs = []
        for (aa, bb, cc) in ((a, b, c), (a2, b2, c + sep)):
            f = aa * (ploty - 1099) ** 2 + bb * (ploty - 1099) + cc
            f = f[(f <= W - 1) & (f >= 0)]
            xs.append(f.astype(int))
        lx, rx = xs
        if len(lx) == 0 or len(rx) == 0:
            continue
        ly = np.arange(H - len(lx), H)
        ry = np.arange(H - len(rx), H)
        canvas = np.zeros((H, W, 3), np.uint8)
        pl = np.array([np.transpose(np.vstack([lx, ly]))])
        pr = np.array([np.flipud(np.transpose(np.vstack([rx, ry])))])
        cv2.fillPoly(canvas, np.int_([np.hstack((pl, pr))]), (0, 255, 0))
        lo, hi = cvops.lane_polygon_rows(lx, ly, rx, ry, W, H)
        assert np.array_equal(cvops.lane_canvas(lo, hi, W, H), canvas), t
        checked += 1
    assert checked > 80


def test_text_sprites_reproduce_puttext():
    """lane_tracker_b200/data glyph sprites == cv2.putText (HERSHEY_SIMPLEX, scale 1, white, thickness 2, LINE_AA),
    the style of lane_tracker.py:653-659 / 668-672, on the strings process() formats and on random ones."""
    from lane_tracker_b200.text import TextSprites
    from oracle.text import overlay_strings, render
    sp = TextSprites.load()
    rng = np.random.default_rng(4)
    cases = overlay_strings(True, 10403, -0.0731520, 971, True) + overlay_strings(False, None, None, 12, True)
    cases += [("".join(chr(c) for c in rng.integers(32, 127, 24)), (int(rng.integers(0, 50)), int(rng.integers(30, 120))))
              for _ in range(40)]
    for text, org in cases:
        bg = rng.integers(0, 256, (140, 700, 3), dtype=np.uint8)
        want = bg.copy()
        cv2.putText(want, text, org, cv2.FONT_HERSHEY_SIMPLEX, fontScale=1, color=(255, 255, 255), thickness=2,
                    lineType=cv2.LINE_AA)
        assert np.array_equal(render(sp, bg.copy(), text, org), want), text
    assert overlay_strings(True, 2784, -0.004, 5, False)[1][0] == "Eccentricity: -0.00 m"


@pytest.mark.parametrize("shape,dsize", [((1100, 1080, 3), (640, 652)), ((1100, 1080), (640, 652)), ((720, 1280, 3), (1280, 720)),
                                         ((300, 500, 3), (777, 333)), ((64, 64, 3), (100, 100)), ((300, 500, 3), (250, 150)),
                                         ((1650, 1620, 3), (960, 978))])
def test_resize_linear(shape, dsize):
    """cv2.resize(img, dsize) of utils.create_split_view (utils.py:88): shrink, enlarge, identity, exact 2x."""
    img = np.random.default_rng(11).integers(0, 256, shape, dtype=np.uint8)
    assert np.array_equal(cvops.resize_linear(img, dsize), cv2.resize(img, dsize=dsize))


@pytest.mark.parametrize("beta", [0.5, 0.3])
def test_add_weighted_general(noise, beta):
    other = np.random.default_rng(12).integers(0, 256, noise.shape, dtype=np.uint8)
    assert np.array_equal(cvops.add_weighted(noise, other, beta), cv2.addWeighted(noise, 1, other, beta, 0))


def test_fill_poly_search_bands_leaving_the_canvas():
    """visualize_band_search polygons (lane_tracker.py:749-758): x -/+ bandwidth may leave the canvas, where
    cv::Line clips an edge before rasterising it."""
    W, H = 1080, 1100
    rng = np.random.default_rng(3)
    checked = clipped = 0
    for t in range(200):
        a = rng.normal(0, 3e-4) * (5 if t % 3 == 0 else 1)
        b = rng.normal(0, 0.5) * (3 if t % 3 == 0 else 1)
        c = rng.uniform(-50, 1130)
        partial = 1.0 if t % 2 == 0 else 0.5
        ploty = np.linspace(H * (1 - partial), H - 1, int(H * partial))
        f = a * (ploty - 1099) ** 2 + b * (ploty - 1099) + c
        x = f[(f <= W - 1) & (f >= 0)].astype(int)
        if len(x) == 0:
            continue
        y = np.arange(H - len(x), H)
        bw = int(rng.choice([25, 40, 80, 120]))
        canvas = np.zeros((H, W, 3), np.uint8)
        w1 = np.array([np.transpose(np.vstack([x - bw, y]))])
        w2 = np.array([np.flipud(np.transpose(np.vstack([x + bw, y])))])
        cv2.fillPoly(canvas, np.int_([np.hstack((w1, w2))]), (0, 255, 0))
        lo, hi = cvops.lane_polygon_rows(x - bw, y, x + bw, y, W, H)
        assert np.array_equal(cvops.lane_canvas(lo, hi, W, H), canvas), t
        checked += 1
        clipped += int((x - bw).min() < 0 or (x + bw).max() > W - 1)
    assert checked > 150 and clipped > 30


@pytest.mark.parametrize("shape", [(720, 1280), (1080, 1920), (16, 64)])
def test_yuv2rgb_nv12(shape):
    """cv2.cvtColor(COLOR_YUV2RGB_NV12): the colour definition the NV12 ingest kernel (lt_nv12_to_rgb) is pinned to.
    Random planes plus every (Y, U, V) extreme, so that the saturation of all three channels is exercised."""
    h, w = shape
    rng = np.random.default_rng(21)
    nv = rng.integers(0, 256, (h * 3 // 2, w), dtype=np.uint8)
    nv[:2, :8] = [[0, 255, 16, 235, 15, 17, 1, 254]] * 2
    nv[h, :8] = [0, 0, 255, 255, 0, 255, 255, 0]
    assert np.array_equal(cvops.yuv2rgb_nv12(nv, w, h), cv2.cvtColor(nv, cv2.COLOR_YUV2RGB_NV12))


def test_lab_b_never_saturates():
    """The 8-bit saturation at the end of OpenCV's RGB -> Lab b (lane_tracker.py:208) never triggers: over all 2^24 RGB
    inputs the value lies well inside [0, 255].  k_warp_planes relies on it (no clamp in its Lab path)."""
    g, cb = (np.asarray(t, dtype=np.int64) for t in cvops.lab_tables())
    B = g[np.arange(256)]
    lo, hi = 1 << 30, -(1 << 30)
    for R in range(256):
        Y = (871 * g[R] + 2929 * g[:, None] + 296 * B[None, :] + 2048) >> 12        # [G, B]
        Z = (73 * g[R] + 448 * g[:, None] + 3575 * B[None, :] + 2048) >> 12
        v = (200 * (cb[Y] - cb[Z]) + 128 * 32768 + 16384) >> 15
        lo, hi = min(lo, int(v.min())), max(hi, int(v.max()))
    assert (lo, hi) == (20, 223)
