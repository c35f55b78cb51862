"""CPU oracle for the lane_tracker per-frame hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline``
/ ``--impl reference`` legs may import it, and only as the checker or as the
reported CPU baseline.  The product path (``lane_tracker_b200``) never imports
this package and fails loudly when its CUDA library is missing.

What is restated, and from where (all citations into /root/reference):

* ``oracle.cvops``   -- NumPy restatements of every third-party (OpenCV 4.13.0,
  un-vendored, un-pinned by the reference: README.md:41-47) operator the path
  calls: lane_tracker.py:832 (undistort), :834/:650 (warpPerspective), :208
  (RGB2LAB), :210-211/:238 (morphologyEx), :73-76 (filter2D), :217-218
  (adaptiveThreshold), :647 (fillPoly), :662 (addWeighted).
* ``oracle.tracker`` -- restatement of ``LaneTracker`` (lane_tracker.py:85-1209)
  and ``bilateral_adaptive_threshold`` (lane_tracker.py:14-83).

Parity pin: the reference ships no tests or golden vectors, so
the oracle is pinned against (a) cv2 4.13.0 itself, op by op, and (b) outputs
of the reference's own ``process()`` imported from /root/reference in the
build container; those outputs are committed under ``tests/golden/`` together
with ``tests/golden/make_golden.py`` which generated them.
"""
