#!/usr/bin/env python3
"""Generate tests/golden/golden.json from the LIVE reference (build container only).

Run:  python tests/golden/make_golden.py

Imports the reference from /root/reference (patched per SURVEY.md Appendix C, see
tests/_liveref.py), runs its own ``process()`` / ``filter_lane_points()`` /
``sliding_window_search()`` on
  * the 11 bundled test frames (copied verbatim to tests/golden/frames/ -- data
    fixtures, not source), each with a fresh tracker, and
  * a 48-frame scenario on one tracker: synthetic road frames (sliding-window then
    band-search tracking), a 12-frame outage (bundled test frame, every attempt
    invalid: exercises n_reset / n_fail), then recovery,
and records sha256 digests of every stage output plus the small numeric results.
The reference itself ships no golden vectors or tests.
"""
import glob
import hashlib
import json
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import _liveref  # noqa: E402
from lane_tracker_b200 import synth  # noqa: E402

TEXT_BOX = (111, 620)  # rows, cols blanked for the legacy 'out' digests; 'out_full' digests cover the whole frame


def sha(a):
    a = np.ascontiguousarray(a)
    return hashlib.sha256(a.tobytes()).hexdigest()[:32]


def out_digest(frame):
    f = frame.copy()
    f[:TEXT_BOX[0], :TEXT_BOX[1]] = 0
    return sha(f)


def load_frame(path):
    import cv2
    return cv2.cvtColor(cv2.imread(path), cv2.COLOR_BGR2RGB)


def pix_digest(t):
    return sha(np.stack([np.asarray(t.left_y), np.asarray(t.left_x)]).astype(np.int64)) + ":" + \
        sha(np.stack([np.asarray(t.right_y), np.asarray(t.right_x)]).astype(np.int64))


def state_record(t):
    def arr(x):
        return None if x is None else [float(v) for v in np.asarray(x).ravel()]
    return dict(
        last_detection=int(t.last_detection), counter=int(t.counter), success=int(t.success),
        detected_pixels=bool(t.detected_pixels), valid=bool(t.valid_lane_lines),
        last_left=arr(t.last_left_coeffs), last_right=arr(t.last_right_coeffs),
        left_avg=arr(t.left_avg_coeffs), right_avg=arr(t.right_avg_coeffs),
        n_left_avg=int(np.asarray(t.left_avg_x).size), n_right_avg=int(np.asarray(t.right_avg_x).size),
        avg_xy=sha(np.concatenate([np.asarray(v, dtype=np.int64).ravel() for v in
                                   (t.left_avg_y, t.left_avg_x, t.right_avg_y, t.right_avg_x)])),
        radius=None if t.average_curve_radius is None else int(t.average_curve_radius),
        radii=[int(v) for v in t.average_curve_radii],
        ecc=None if t.eccentricity is None else float(t.eccentricity),
        n_left=None if t.left_x is None else int(len(t.left_x)),
        n_right=None if t.right_x is None else int(len(t.right_x)),
        pix=None if t.left_x is None else pix_digest(t))


def stage_record(ref, frame):
    """Per-stage digests using the reference's own calls on one frame (fresh tracker)."""
    import cv2
    und = cv2.undistort(frame, ref.cam_matrix, ref.dist_coeffs, None, ref.cam_matrix)
    bv = cv2.warpPerspective(und, ref.M, ref.warped_size, flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT)
    rec = dict(frame=sha(frame), undistorted_rows_457_694=sha(und[457:695]), bv=sha(bv))
    m1 = ref.filter_lane_points(bv, filter_type="bilateral", ksize_r=15, C_r=8, ksize_b=35, C_b=5,
                                mask_noise=False, ksize_noise=65, C_noise=10, noise_thresh=140)
    m2 = ref.filter_lane_points(bv, filter_type="neighborhood", ksize_r=15, C_r=5, ksize_b=35, C_b=5,
                                mask_noise=False, ksize_noise=65, C_noise=10, noise_thresh=140)
    m3 = ref.filter_lane_points(bv, filter_type="bilateral", ksize_r=15, C_r=8, ksize_b=35, C_b=5,
                                mask_noise=True, ksize_noise=65, C_noise=10, noise_thresh=140)
    rec.update(mask_bilateral=sha(m1), mask_neighborhood=sha(m2), mask_bilateral_noise=sha(m3),
               mask_counts=[int((m1 > 0).sum()), int((m2 > 0).sum()), int((m3 > 0).sum())])
    for name, m, nsl in (("sws_bilateral", m1, 8), ("sws_neighborhood", m2, 50)):
        ref.detected_pixels = False
        ref.sliding_window_search(m, window_width=30, window_height=40, search_range=20, mu=0.1,
                                  no_success_limit=nsl, start_slice=0.25, ignore_sides=360,
                                  ignore_bottom=30, partial=1.0)
        r = dict(detected=bool(ref.detected_pixels))
        if ref.detected_pixels:
            lf, rf = ref.fit_poly()
            ref.check_validity(lf, rf)
            r.update(pix=pix_digest(ref), n_left=int(len(ref.left_x)), n_right=int(len(ref.right_x)),
                     left_centroids=[int(v) for v in ref.left_window_centroids],
                     right_centroids=[int(v) for v in ref.right_window_centroids],
                     left_fit=[float(v) for v in lf], right_fit=[float(v) for v in rf],
                     valid=bool(ref.valid_lane_lines))
        rec[name] = r
    return rec


def debug_views_record():
    """Reference debug views (visualize_search / split_view, lane_tracker.py:689-793, 1130-1209) on an 8-frame
    sequence: sliding-window frame, band-search frames, a two-frame outage (attempt 2, still pixels) and recovery."""
    outage = load_frame(os.path.join(HERE, "frames", "test2.jpg"))
    vid = synth.RoadVideo(1)
    frames = [("synth", t) for t in range(5)] + [("outage", 5), ("outage", 6), ("synth", 7)]
    rec = dict(seed=1, outage_frame="test2.jpg", frames=[[k, t] for k, t in frames], visualize_search=[], split_view=[])
    for mode in ("visualize_search", "split_view"):
        ref = _liveref.make_tracker()
        for kind, t in frames:
            f = outage if kind == "outage" else vid.frame(t)
            r = _liveref.quiet(ref.process, f.copy(), **{mode: True})
            if mode == "visualize_search":
                rec[mode].append(dict(out=sha(r[0]), vis=sha(r[1]), vis_shape=list(r[1].shape)))
            else:
                rec[mode].append(dict(canvas=sha(r), shape=list(r.shape)))
            print(mode, kind, t, rec[mode][-1])
    return rec


def main():
    warnings.simplefilter("ignore")
    import cv2
    if "--debug-views-only" in sys.argv:            # refresh one section, keep the rest of the file
        with open(os.path.join(HERE, "golden.json")) as f:
            gold = json.load(f)
        gold["debug_views"] = debug_views_record()
        with open(os.path.join(HERE, "golden.json"), "w") as f:
            json.dump(gold, f, indent=1)
        print("updated golden.json: debug_views")
        return
    gold = dict(meta=dict(cv2=cv2.__version__, numpy=np.__version__, text_box=list(TEXT_BOX),
                          generator="tests/golden/make_golden.py", reference="pierluigiferrari/lane_tracker"))
    gold["images"] = {}
    for path in sorted(glob.glob(os.path.join(HERE, "frames", "*.jpg"))):
        name = os.path.basename(path)
        frame = load_frame(path)
        rec = stage_record(_liveref.make_tracker(), frame)
        ref = _liveref.make_tracker()
        out = _liveref.quiet(ref.process, frame.copy())
        rec["process_out"] = out_digest(out)
        rec["process_out_full"] = sha(out)
        rec["process_state"] = state_record(ref)
        gold["images"][name] = rec
        print(name, rec["mask_counts"], rec["process_state"]["valid"])
    # scenario
    outage = load_frame(os.path.join(HERE, "frames", "test2.jpg"))
    vid = synth.RoadVideo(0)
    ref = _liveref.make_tracker()
    seq = []
    for t in range(48):
        kind = "outage" if 14 <= t < 26 else "synth"
        frame = outage if kind == "outage" else vid.frame(t)
        out = _liveref.quiet(ref.process, frame.copy())
        rec = dict(t=t, kind=kind, frame=sha(frame), out=out_digest(out), out_full=sha(out), state=state_record(ref))
        seq.append(rec)
        print(t, kind, rec["state"]["valid"], rec["state"]["last_detection"], rec["state"]["radius"])
    gold["scenario"] = dict(seed=0, frames=seq, outage_frame="test2.jpg", outage=[14, 26])
    ratio = ref.get_success_ratio()
    gold["scenario"]["success_ratio"] = [float(ratio[0]), int(ratio[1]), int(ratio[2])]
    gold["debug_views"] = debug_views_record()
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(gold, f, indent=1)
    print("wrote golden.json")


if __name__ == "__main__":
    main()
