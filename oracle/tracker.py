"""CPU restatement of the reference ``LaneTracker`` (lane_tracker.py:85-1209).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

The search / fit / validity / state-machine logic is restated in integer form
 rather than transcribed; the image operators come
from ``oracle.cvops`` (NumPy restatements, default) or, with ``backend='cv2'``,
from the same cv2 calls the reference makes (used for the timed CPU baseline,
where the NumPy restatements would be an unfairly slow stand-in).

``cv2.putText`` overlays (lane_tracker.py:653-659, 668-672): drawn with cv2 itself under
``backend='cv2'`` and with the glyph sprites of ``lane_tracker_b200/data`` (generated from, and checked
bit-for-bit against, cv2.putText by tools/make_text_sprites.py) under ``backend='numpy'``;
``render_text=False`` skips them.  Unlike the reference the caller's frame is not modified.
Not restated: the debug views (675-793).
"""
from __future__ import annotations

import math

import numpy as np

from . import cvops

# process() keyword defaults, lane_tracker.py:876-900
PROCESS_DEFAULTS = dict(
    ksize_r=15, C_r=8, ksize_b=35, C_b=5, filter_type="bilateral", mask_noise=False,
    noise_thresh=140, ksize_noise=65, C_noise=10, window_width=30, window_height=40,
    search_range=20, mu=0.1, no_success_limit=8, start_slice=0.25, ignore_sides=360,
    ignore_bottom=30, bandwidth=25, partial=1.0, n_tries=2)

# hard-coded second attempt, lane_tracker.py:1081-1099
ATTEMPT2 = dict(
    ksize_r=15, C_r=5, ksize_b=35, C_b=5, filter_type="neighborhood", mask_noise=False,
    noise_thresh=140, ksize_noise=65, C_noise=10, window_width=30, window_height=40,
    search_range=20, mu=0.1, no_success_limit=50, start_slice=0.25, ignore_sides=360,
    ignore_bottom=30, bandwidth=30, partial=1.0)

# check_validity constants, lane_tracker.py:588-593, 617
VALIDITY = dict(min_d1=150, max_d1=230, min_d2=110, max_d2=230, min_d3=80, max_d3=200, tan=0.25)


def _pyslice(start, stop, n):
    """Resolve ``a[start:stop]`` on a length-n axis the way Python does."""
    if start < 0:
        start = max(start + n, 0)
    else:
        start = min(start, n)
    if stop < 0:
        stop = max(stop + n, 0)
    else:
        stop = min(stop, n)
    return start, max(stop, start)


def window_sums(col, width):
    """``np.convolve(np.ones(width), col)`` ('full') on integer counts."""
    n = len(col)
    cs = np.concatenate([[0], np.cumsum(col, dtype=np.int64)])
    i = np.arange(n + width - 1)
    hi = np.minimum(i, n - 1) + 1
    lo = np.maximum(i - (width - 1), 0)
    return cs[hi] - cs[lo]


class OracleLaneTracker:
    """Same constructor and methods as the reference class (lane_tracker.py:101)."""

    def __init__(self, img_size, warped_size, cam_matrix, dist_coeffs, warp_matrices,
                 mpp_conversion, n_fail=8, n_reset=4, n_average=2, print_frame_count=False,
                 backend="numpy", render_text=True):
        self.img_size = tuple(img_size)
        self.warped_size = tuple(warped_size)
        self.cam_matrix = np.asarray(cam_matrix, dtype=np.float64)
        self.dist_coeffs = np.asarray(dist_coeffs, dtype=np.float64)
        self.M = np.asarray(warp_matrices[0], dtype=np.float64)
        self.Minv = np.asarray(warp_matrices[1], dtype=np.float64)
        self.mppv, self.mpph = mpp_conversion
        self.n_fail, self.n_reset, self.n_average = n_fail, n_reset, n_average
        self.print_frame_count = print_frame_count
        self.backend = backend
        self.render_text = render_text
        self._sprites = None
        self.validity = dict(VALIDITY)    # per-instance copy: tests may install the other documented windows
        if backend == "cv2":
            import cv2  # noqa: F401  (third-party kernel library of the reference)
            self._cv2 = cv2
        # state, lane_tracker.py:139-176
        self.last_detection = n_reset + 1
        self.detected_pixels = False
        self.valid_lane_lines = False
        self.left_fit_coeffs, self.right_fit_coeffs = [], []
        self.last_left_coeffs = self.last_right_coeffs = None
        self.left_avg_coeffs = self.right_avg_coeffs = None
        self.left_avg_y = np.array([])
        self.left_avg_x = np.array([])
        self.right_avg_y = np.array([])
        self.right_avg_x = np.array([])
        self.left_y = self.left_x = self.right_y = self.right_x = None
        self.left_window_centroids = self.right_window_centroids = None
        self.left_curve_radius = self.right_curve_radius = None
        self.average_curve_radius = None
        self.average_curve_radii = []
        self.eccentricity = None
        self.counter = 0
        self.success = 0
        self.trace = {}
        self._maps = None

    # ------------------------------------------------------------------ ops
    def _remap(self, img):
        """undistort + warp to the bird's-eye view (lane_tracker.py:832-834)."""
        if self.backend == "cv2":
            cv2 = self._cv2
            u = cv2.undistort(img, self.cam_matrix, self.dist_coeffs, None, self.cam_matrix)
            return cv2.warpPerspective(u, self.M, self.warped_size, flags=cv2.INTER_LINEAR,
                                       borderMode=cv2.BORDER_CONSTANT)
        if self._maps is None:
            w, h = self.img_size
            self._maps = (cvops.undistort_map_q5(self.cam_matrix, self.dist_coeffs, w, h),
                          cvops.perspective_map_q5(self.M, *self.warped_size))
        (U, V), (X, Y) = self._maps
        und = cvops.bilinear_q5(img, U, V)
        self.trace["undistorted"] = und
        return cvops.bilinear_q5(und, X, Y)

    def get_success_ratio(self):
        return self.success / self.counter, self.success, self.counter

    def filter_lane_points(self, img, filter_type="bilateral", ksize_r=25, C_r=8, ksize_b=35,
                           C_b=5, mask_noise=False, ksize_noise=65, C_noise=10, noise_thresh=135):
        """lane_tracker.py:183-240."""
        if filter_type not in ("bilateral", "neighborhood"):
            raise ValueError("Unexpected filter mode. Expected modes are 'bilateral' or 'neighborhood'.")
        tr = self.trace
        r_plane = np.ascontiguousarray(img[:, :, 0])
        if self.backend == "cv2":
            cv2 = self._cv2
            b_plane = cv2.cvtColor(img, cv2.COLOR_RGB2LAB)[:, :, 2]
            if filter_type == "bilateral":
                se_b = cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (55, 55))
                se_r = cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (29, 29))
                r_top = cv2.morphologyEx(r_plane, cv2.MORPH_TOPHAT, se_r)
                b_top = cv2.morphologyEx(b_plane, cv2.MORPH_TOPHAT, se_b)
                r_thr = self._cross_cv2(r_top, ksize_r, C_r)
                b_thr = self._cross_cv2(b_top, ksize_b, C_b)
            else:
                r_thr = cv2.adaptiveThreshold(r_plane, 255, cv2.ADAPTIVE_THRESH_MEAN_C,
                                              cv2.THRESH_BINARY, ksize_r, -C_r)
                b_thr = cv2.adaptiveThreshold(b_plane, 255, cv2.ADAPTIVE_THRESH_MEAN_C,
                                              cv2.THRESH_BINARY, ksize_b, -C_b)
            merged = (r_thr > 0) | (b_thr > 0)
            if mask_noise:
                n1 = cv2.inRange(b_plane, noise_thresh, 255)
                n2 = self._cross_cv2(b_plane, ksize_noise, C_noise)
                merged &= (n1 == 0) | (n2 > 0)
            merged = np.where(merged, 255, 0).astype(np.uint8)
            se_o = cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (5, 5))
            return cv2.morphologyEx(merged, cv2.MORPH_OPEN, se_o)
        b_plane = cvops.lab_b_plane(img)
        tr["r_plane"], tr["b_plane"] = r_plane, b_plane
        if filter_type == "bilateral":
            # the reference also computes (and discards) both top-hats in
            # 'neighborhood' mode (:210-211); the result does not depend on it.
            r_top = cvops.tophat_ellipse(r_plane, 29)
            b_top = cvops.tophat_ellipse(b_plane, 55)
            tr["r_tophat"], tr["b_tophat"] = r_top, b_top
            r_thr = cvops.cross_threshold(r_top, ksize_r, C_r)
            b_thr = cvops.cross_threshold(b_top, ksize_b, C_b)
        else:
            r_thr = cvops.box_mean_threshold(r_plane, ksize_r, C_r)
            b_thr = cvops.box_mean_threshold(b_plane, ksize_b, C_b)
        tr["r_thresh"], tr["b_thresh"] = r_thr, b_thr
        merged = (r_thr > 0) | (b_thr > 0)
        if mask_noise:
            n1 = cvops.in_range(b_plane, noise_thresh, 255)
            n2 = cvops.cross_threshold(b_plane, ksize_noise, C_noise)
            merged &= (n1 == 0) | (n2 > 0)
        merged = np.where(merged, 255, 0).astype(np.uint8)
        tr["merged"] = merged
        return cvops.open_ellipse(merged, 5)

    def _cross_cv2(self, img, k, C):
        cv2 = self._cv2
        kl = np.array([[1] * k + [-k]], dtype=np.int16)
        kr = np.array([[-k] + [1] * k], dtype=np.int16)
        d = C * k
        a = cv2.filter2D(img, cv2.CV_16S, kl, anchor=(k, 0), delta=d, borderType=cv2.BORDER_CONSTANT)
        b = cv2.filter2D(img, cv2.CV_16S, kr, anchor=(0, 0), delta=d, borderType=cv2.BORDER_CONSTANT)
        c = cv2.filter2D(img, cv2.CV_16S, kl.T.copy(), anchor=(0, k), delta=d, borderType=cv2.BORDER_CONSTANT)
        e = cv2.filter2D(img, cv2.CV_16S, kr.T.copy(), anchor=(0, 0), delta=d, borderType=cv2.BORDER_CONSTANT)
        return np.where(((a < 0) & (b < 0)) | ((c < 0) & (e < 0)), 255, 0).astype(np.uint8)

    # --------------------------------------------------------------- search
    def sliding_window_search(self, img, window_width, window_height, search_range, mu,
                              no_success_limit, start_slice=0.25, ignore_sides=360,
                              ignore_bottom=30, partial=1, diagnostics=False):
        """lane_tracker.py:242-447, integer restatement."""
        B = img > 0
        W = img.shape[1]
        Hh = img.shape[0] - ignore_bottom
        cx = int(W / 2)
        y0 = int((1 - start_slice) * Hh)
        hw = int(window_width / 2)
        nlev = int((partial * Hh) / window_height)
        sides = []
        hist0 = []
        for (lo, hi, dflt) in ((ignore_sides, cx, int(W * 0.4)), (cx, W - ignore_sides, int(W * 0.6))):
            r0, r1 = _pyslice(y0, Hh, B.shape[0])
            c0, c1 = _pyslice(lo, hi, W)
            col = B[r0:r1, c0:c1].sum(axis=0)
            hist0.append(col)
            st = dict(c=dflt, cents=[], diffs=[], miss=0, rmin=-search_range, rmax=search_range,
                      ys=[], xs=[], hit0=False)
            if col.any():
                S = window_sums(col, window_width)
                idx = np.nonzero(S == S.max())[0]
                st["c"] = int((idx[0] + idx[-1]) / 2) - hw + lo
                self._collect(B, st, Hh - window_height, Hh, hw)
                st["hit0"] = True
            st["cents"].append(st["c"])
            sides.append(st)
        self.trace["sws_hist0"] = hist0
        level_S = []
        for level in range(1, nlev):
            ra, rb = Hh - (1 + level) * window_height, Hh - level * window_height
            r0, r1 = _pyslice(ra, rb, B.shape[0])
            S = window_sums(B[r0:r1].sum(axis=0), window_width)
            level_S.append(S)
            for s, st in enumerate(sides):
                other = sides[1 - s]
                if st["miss"] >= no_success_limit:
                    continue
                lo = max(st["c"] + st["rmin"] + hw, 0)
                hi = min(st["c"] + st["rmax"] + hw, W)
                a, b = _pyslice(lo, hi, len(S))
                seg = S[a:b]
                if seg.size and seg.max() > 0:
                    idx = np.nonzero(seg == seg.max())[0]
                    mc = int(math.ceil((idx[0] + idx[-1]) / 2))
                    st["c"] = mc + lo - hw
                    st["cents"].append(st["c"])
                    st["diffs"].append(st["cents"][-1] - st["cents"][-2])
                    st["miss"] = 0
                    self._collect(B, st, ra, rb, hw)
                    d = int(mu * st["diffs"][-1])
                    st["rmin"] += d
                    st["rmax"] += d
                else:
                    if len(other["diffs"]) > 0 and other["miss"] == 0:
                        st["c"] += int(other["diffs"][-1])
                    st["cents"].append(st["c"])
                    st["miss"] += 1
                    if st["miss"] >= no_success_limit:
                        del st["cents"][-no_success_limit:]
        self.trace["sws_level_S"] = level_S
        L, R = sides
        nl = sum(len(a) for a in L["xs"])
        nr = sum(len(a) for a in R["xs"])
        if len(L["xs"]) > 0 and len(R["xs"]) > 0 and nl > 0 and nr > 0:
            self.left_y = np.concatenate(L["ys"])
            self.left_x = np.concatenate(L["xs"])
            self.right_y = np.concatenate(R["ys"])
            self.right_x = np.concatenate(R["xs"])
            self.detected_pixels = True
            self.left_window_centroids = L["cents"]
            self.right_window_centroids = R["cents"]
        else:
            self.detected_pixels = False
        self.trace["sws_centroids"] = (list(L["cents"]), list(R["cents"]))

    @staticmethod
    def _collect(B, st, ra, rb, hw):
        """nonzero pixels of ``img[ra:rb, c-hw:c+hw]`` with NumPy slice semantics."""
        c = st["c"]
        r0, r1 = _pyslice(ra, rb, B.shape[0])
        c0, c1 = _pyslice(c - hw, c + hw, B.shape[1])
        ys, xs = np.nonzero(B[r0:r1, c0:c1])
        st["ys"].append(ys + ra)
        st["xs"].append(xs + (c - hw))

    def band_search(self, img, bandwidth, ignore_bottom=30, partial=1, diagnostics=False):
        """lane_tracker.py:449-500."""
        B = img > 0
        Hh = img.shape[0]
        B = B.copy()
        B[Hh - ignore_bottom:, :] = False
        B[:int(Hh * (1 - partial)), :] = False
        ys, xs = np.nonzero(B)
        sel = []
        for cf in (self.last_left_coeffs, self.last_right_coeffs):
            f = cf[0] * (ys ** 2) + cf[1] * ys + cf[2]
            sel.append((xs > f - bandwidth) & (xs < f + bandwidth))
        if sel[0].any() and sel[1].any():
            self.left_y, self.left_x = ys[sel[0]], xs[sel[0]]
            self.right_y, self.right_x = ys[sel[1]], xs[sel[1]]
            self.detected_pixels = True
        else:
            self.detected_pixels = False

    # ------------------------------------------------------------------ fit
    def fit_poly(self):
        """lane_tracker.py:502-509."""
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            return (np.polyfit(self.left_y, self.left_x, 2), np.polyfit(self.right_y, self.right_x, 2))

    def get_poly_points(self, left_fit_coeffs, right_fit_coeffs, partial=1):
        """lane_tracker.py:511-528 (with the NumPy<=1.11 truncations)."""
        Wd, Hh = self.warped_size
        ploty = np.linspace(Hh * (1 - partial), Hh - 1, int(Hh * partial))
        out = []
        for cf in (left_fit_coeffs, right_fit_coeffs):
            fx = cf[0] * ploty ** 2 + cf[1] * ploty + cf[2]
            keep = fx[(fx <= Wd - 1) & (fx >= 0)]
            fy = np.linspace(Hh - len(keep), Hh - 1, len(keep))
            out += [fy.astype(int), keep.astype(int)]
        return tuple(out)

    def check_validity(self, left_fit_coeffs, right_fit_coeffs, diagnostics=False):
        """lane_tracker.py:561-627."""
        ly, _, ry, _ = self.get_poly_points(left_fit_coeffs, right_fit_coeffs)
        Wd = self.warped_size[0]
        n = min(len(ly), len(ry))
        y1 = Wd - 1
        y2 = Wd - int(n * 0.35)
        y3 = Wd - int(n * 0.75)
        l, r = left_fit_coeffs, right_fit_coeffs

        def f(c, y):
            return c[0] * (y ** 2) + c[1] * y + c[2]
        d1, d2, d3 = abs(f(l, y1) - f(r, y1)), abs(f(l, y2) - f(r, y2)), abs(f(l, y3) - f(r, y3))
        V = self.validity
        self.trace["validity"] = (d1, d2, d3)
        if (d1 < V["min_d1"]) | (d1 > V["max_d1"]) | (d2 < V["min_d2"]) | (d2 > V["max_d2"]) | \
                (d3 < V["min_d3"]) | (d3 > V["max_d3"]):
            self.valid_lane_lines = False
            return
        n1 = abs((2 * l[0] * y1 + l[1]) - (2 * r[0] * y1 + r[1]))
        n2 = abs((2 * l[0] * y3 + l[1]) - (2 * r[0] * y3 + r[1]))
        self.valid_lane_lines = not ((n1 >= V["tan"]) | (n2 >= V["tan"]))

    def get_curve_radius(self):
        """lane_tracker.py:530-549."""
        lm = np.polyfit(self.left_y * self.mppv, self.left_x * self.mpph, 2)
        rm = np.polyfit(self.right_y * self.mppv, self.right_x * self.mpph, 2)
        ye = self.warped_size[1]
        self.left_curve_radius = int(((1 + (2 * lm[0] * ye * self.mppv + lm[1]) ** 2) ** 1.5) / np.absolute(2 * lm[0]))
        self.right_curve_radius = int(((1 + (2 * rm[0] * ye * self.mppv + rm[1]) ** 2) ** 1.5) / np.absolute(2 * rm[0]))
        self.average_curve_radii.append(int(0.5 * (self.left_curve_radius + self.right_curve_radius)))
        if len(self.average_curve_radii) > self.n_average:
            self.average_curve_radii.pop(0)
        self.average_curve_radius = int(np.average([r for r in self.average_curve_radii if r > 0]))

    def get_eccentricity(self):
        """lane_tracker.py:551-559."""
        mid = int(self.warped_size[0] / 2)
        self.eccentricity = (((mid - self.left_avg_x[-1]) - (self.right_avg_x[-1] - mid)) / 2) * self.mpph

    def _put_text(self, img, drew_lane):
        """The putText calls of draw_lane (:653-659) / print_failure (:668-672), on a copy of the frame."""
        if not self.render_text:
            return img
        from lane_tracker_b200.text import TextSprites
        from oracle.text import overlay_strings, render
        img = img.copy()
        for text, org in overlay_strings(drew_lane, self.average_curve_radius, self.eccentricity, self.counter,
                                         self.print_frame_count):
            if self.backend == "cv2":
                cv2 = self._cv2
                cv2.putText(img, text, org, cv2.FONT_HERSHEY_SIMPLEX, fontScale=1, color=(255, 255, 255),
                            thickness=2, lineType=cv2.LINE_AA)
            else:
                if self._sprites is None:
                    self._sprites = TextSprites.load()
                render(self._sprites, img, text, org)
        return img

    def draw_lane(self, img):
        """lane_tracker.py:629-662."""
        Wd, Hh = self.warped_size
        lo, hi = cvops.lane_polygon_rows(self.left_avg_x, self.left_avg_y, self.right_avg_x,
                                         self.right_avg_y, Wd, Hh)
        self.trace["lane_rows"] = (lo, hi)
        canvas = cvops.lane_canvas(lo, hi, Wd, Hh)
        img = self._put_text(img, True)
        if self.backend == "cv2":
            cv2 = self._cv2
            unwarped = cv2.warpPerspective(canvas, self.Minv, (img.shape[1], img.shape[0]))
            return cv2.addWeighted(img, 1, unwarped, 0.3, 0)
        unwarped = cvops.warp_perspective(canvas, self.Minv, (img.shape[1], img.shape[0]))
        return cvops.add_weighted_03(img, unwarped)

    def print_failure(self, img):
        """lane_tracker.py:664-673."""
        return self._put_text(img, False)

    # ---------------------------------------------------------- debug views
    def window_rect(self, shape, window_width, window_height, center, level, ignore_bottom):
        """Rows/columns set by ``window_mask`` (lane_tracker.py:675-687), Python slice semantics included."""
        H, W = shape
        img_height = H - ignore_bottom
        r0, r1 = _pyslice(int(img_height - (level + 1) * window_height), int(img_height - level * window_height), H)
        c0, c1 = _pyslice(max(int(center - window_width / 2), 0), min(int(center + window_width / 2), W), W)
        return r0, r1, c0, c1

    def visualize_sliding_window_search(self, binary_img, left_fit_coeffs, right_fit_coeffs, window_width,
                                        window_height, ignore_bottom):
        """lane_tracker.py:689-729."""
        pts = []
        for cents in (self.left_window_centroids, self.right_window_centroids):
            m = np.zeros_like(binary_img)
            for level, c in enumerate(cents):
                r0, r1, c0, c1 = self.window_rect(binary_img.shape, window_width, window_height, c, level, ignore_bottom)
                m[r0:r1, c0:c1] = 255
            pts.append(m)
        template = np.array(pts[1] + pts[0], np.uint8)          # uint8 sum: 255 + 255 wraps to 254
        zero = np.zeros_like(template)
        color = np.stack([binary_img] * 3, axis=-1)
        tmpl3 = np.stack([zero, template, zero], axis=-1)
        if self.backend == "cv2":
            output = self._cv2.addWeighted(color, 1, tmpl3, 0.5, 0.0)
        else:
            output = cvops.add_weighted(color, tmpl3, 0.5)
        output[self.left_y, self.left_x] = [255, 0, 0]
        output[self.right_y, self.right_x] = [0, 0, 255]
        ly, lx, ry, rx = self.get_poly_points(left_fit_coeffs, right_fit_coeffs)
        output[ly, lx] = [255, 235, 0]
        output[ry, rx] = [255, 235, 0]
        return output

    def visualize_band_search(self, binary_img, left_fit_coeffs, right_fit_coeffs, bandwidth, partial):
        """lane_tracker.py:731-771."""
        H, W = binary_img.shape
        output = np.stack([binary_img] * 3, axis=-1)
        window_img = np.zeros_like(output)
        output[self.left_y, self.left_x] = [255, 0, 0]
        output[self.right_y, self.right_x] = [0, 0, 255]
        lby, lbx, rby, rbx = self.get_poly_points(self.last_left_coeffs, self.last_right_coeffs, partial)
        if self.backend == "cv2":
            cv2 = self._cv2
            for bx, by in ((lbx, lby), (rbx, rby)):
                w1 = np.array([np.transpose(np.vstack([bx - bandwidth, by]))])
                w2 = np.array([np.flipud(np.transpose(np.vstack([bx + bandwidth, by])))])
                cv2.fillPoly(window_img, np.int_([np.hstack((w1, w2))]), (0, 255, 0))
            result = cv2.addWeighted(output, 1, window_img, 0.3, 0)
        else:
            for bx, by in ((lbx, lby), (rbx, rby)):
                xl = np.int_(bx - bandwidth)
                xr = np.int_(bx + bandwidth)
                lo, hi = cvops.lane_polygon_rows(xl, by, xr, by, W, H)
                window_img |= cvops.lane_canvas(lo, hi, W, H)
            result = cvops.add_weighted(output, window_img, 0.3)
        ly, lx, ry, rx = self.get_poly_points(left_fit_coeffs, right_fit_coeffs)
        result[ly, lx] = [255, 235, 0]
        result[ry, rx] = [255, 235, 0]
        return result

    def create_split_view(self, target_size, images, positions, sizes):
        """utils.py:57-103 without captions (the tracker passes none)."""
        x_max, y_max = target_size
        canvas = np.zeros((y_max, x_max, 3), dtype=np.uint8)
        for i, img in enumerate(images):
            # the reference's condition, operator precedence included: a != (b | c) != d
            if img.shape[0] != sizes[i][1] | img.shape[1] != sizes[i][0]:
                img = self._cv2.resize(img, dsize=sizes[i]) if self.backend == "cv2" else cvops.resize_linear(img, sizes[i])
            x, y = positions[i]
            w, h = sizes[i]
            canvas[y:min(y + h, y_max), x:min(x + w, x_max), :] = img[:min(h, y_max - y), :min(w, x_max - x)]
        return canvas

    def triple_split_view(self, images):
        """lane_tracker.py:773-793."""
        img1_size = (images[0].shape[1], images[0].shape[0])
        img2_size = (images[1].shape[1], images[1].shape[0])
        positions = [(0, 0), (0, img1_size[1]), (round(0.5 * img1_size[0]), img1_size[1])]
        scale_factor = img2_size[0] / (0.5 * img1_size[0])
        scaled_size = (round(img2_size[0] / scale_factor), round(img2_size[1] / scale_factor))
        target_size = (img1_size[0], img1_size[1] + scaled_size[1])
        return self.create_split_view(target_size, images, positions, [img1_size, scaled_size, scaled_size])

    # -------------------------------------------------------------- process
    def find_lane_points(self, img, **kw):
        """lane_tracker.py:795-874."""
        p = dict(PROCESS_DEFAULTS, mask_noise=True, bandwidth=30, partial=0.5)
        p.update(kw)
        bv = self._remap(img)
        self.trace["bv"] = bv
        mask = self.filter_lane_points(
            bv, filter_type=p["filter_type"], ksize_r=p["ksize_r"], C_r=p["C_r"], ksize_b=p["ksize_b"],
            C_b=p["C_b"], mask_noise=p["mask_noise"], ksize_noise=p["ksize_noise"],
            C_noise=p["C_noise"], noise_thresh=p["noise_thresh"])
        if self.last_detection > self.n_reset:
            self.sliding_window_search(
                mask, window_width=p["window_width"], window_height=p["window_height"],
                search_range=p["search_range"], mu=p["mu"], no_success_limit=p["no_success_limit"],
                start_slice=p["start_slice"], ignore_sides=p["ignore_sides"],
                ignore_bottom=p["ignore_bottom"], partial=p["partial"])
            return mask, "sws"
        self.band_search(mask, bandwidth=p["bandwidth"], ignore_bottom=p["ignore_bottom"],
                         partial=p["partial"])
        return mask, "bs"

    def process(self, img, **kw):
        """lane_tracker.py:876-1209, debug views (``visualize_search`` / ``split_view``) included."""
        p = dict(PROCESS_DEFAULTS)
        p.update(kw)
        n_tries = p.pop("n_tries")
        visualize_search = p.pop("visualize_search", False)
        split_view = p.pop("split_view", False)
        warped_img = None
        if visualize_search | split_view:
            # lane_tracker.py:1035: bird's-eye view of the RAW frame (not undistorted), before any text is drawn
            if self.backend == "cv2":
                warped_img = self._cv2.warpPerspective(img, self.M, self.warped_size, flags=self._cv2.INTER_LINEAR,
                                                       borderMode=self._cv2.BORDER_CONSTANT)
            else:
                warped_img = cvops.warp_perspective(img, self.M, self.warped_size)
        self.counter += 1
        self.detected_pixels = False
        self.valid_lane_lines = False
        self.trace = {"attempts": []}
        lf = rf = None
        for attempt in (1, 2):
            if attempt == 2:
                if not (((not self.detected_pixels) | (not self.valid_lane_lines)) and
                        ((n_tries >= 2) | (n_tries == -1))):
                    break
                p = dict(ATTEMPT2)
            mask, mode = self.find_lane_points(img, **p)
            rec = dict(mode=mode, mask=mask, detected=self.detected_pixels)
            if self.detected_pixels:
                lf, rf = self.fit_poly()
                self.check_validity(lf, rf)
                rec.update(left_fit=lf, right_fit=rf, left_x=self.left_x, left_y=self.left_y,
                           right_x=self.right_x, right_y=self.right_y)
            rec["valid"] = self.valid_lane_lines
            rec["bv"] = self.trace.get("bv")
            self.trace["attempts"].append(rec)
        vis = None
        if visualize_search | split_view:                       # lane_tracker.py:1130-1138
            if self.detected_pixels:
                if mode == "sws":
                    vis = self.visualize_sliding_window_search(mask, lf, rf, p["window_width"], p["window_height"],
                                                               p["ignore_bottom"])
                else:
                    vis = self.visualize_band_search(mask, lf, rf, p["bandwidth"], p["partial"])
            else:
                vis = mask
            self.trace["search_visualization"] = vis
            self.trace["warped_img"] = warped_img

        def finish(out):
            if visualize_search:
                return out, vis
            if split_view:
                return self.triple_split_view([out, warped_img, vis])
            return out

        if not self.valid_lane_lines:
            self.left_fit_coeffs.append(np.array([]))
            self.right_fit_coeffs.append(np.array([]))
            self.average_curve_radii.append(-1)
            if len(self.left_fit_coeffs) > self.n_average:
                self.left_fit_coeffs.pop(0)
                self.right_fit_coeffs.pop(0)
            if len(self.average_curve_radii) > self.n_average:
                self.average_curve_radii.pop(0)
            self.last_detection += 1
            if (self.left_avg_y.size != 0) & (self.last_detection <= self.n_fail):
                return finish(self.draw_lane(img))
            return finish(self.print_failure(img))
        self.left_fit_coeffs.append(lf)
        self.right_fit_coeffs.append(rf)
        self.last_left_coeffs, self.last_right_coeffs = lf, rf
        if len(self.left_fit_coeffs) > self.n_average:
            self.left_fit_coeffs.pop(0)
            self.right_fit_coeffs.pop(0)
        self.last_detection = 0
        self.success += 1
        self.left_avg_coeffs = np.average([c for c in self.left_fit_coeffs if c.size != 0], axis=0)
        self.right_avg_coeffs = np.average([c for c in self.right_fit_coeffs if c.size != 0], axis=0)
        (self.left_avg_y, self.left_avg_x, self.right_avg_y, self.right_avg_x) = \
            self.get_poly_points(self.left_avg_coeffs, self.right_avg_coeffs, p["partial"])
        self.get_curve_radius()
        self.get_eccentricity()
        return finish(self.draw_lane(img))
