"""Parameter sets the reference documents per demo video (tracker_settings.md:1-111).

Each preset has the `process()` keyword arguments and the `check_validity()` windows; the reference ships the
Demo-2 windows as hard-coded constants (lane_tracker.py:588-593) and tells users to edit the source for the others.

    lt = LaneTracker(...); lt.set_validity(**DEMO_1["validity"]); out = lt.process(frame, **DEMO_1["process"])
"""

_COMMON = dict(ksize_r=15, C_r=8, ksize_b=35, C_b=5, filter_type="bilateral", mask_noise=True, noise_thresh=140,
               ksize_noise=65, C_noise=10, window_width=30, window_height=40, search_range=20, mu=0.1,
               no_success_limit=50, start_slice=0.25, ignore_sides=360, ignore_bottom=30, bandwidth=30, partial=1.0,
               n_tries=2)

DEMO_1 = dict(process=dict(_COMMON),
              validity=dict(min_dist_y1=150, max_dist_y1=245, min_dist_y2=150, max_dist_y2=255,
                            min_dist_y3=150, max_dist_y3=255, tangent_thresh=0.25))
DEMO_2 = dict(process=dict(_COMMON, ksize_r=20, C_r=5, mask_noise=False, n_tries=1),
              validity=dict(min_dist_y1=150, max_dist_y1=230, min_dist_y2=110, max_dist_y2=230,
                            min_dist_y3=80, max_dist_y3=200, tangent_thresh=0.25))
DEMO_3 = dict(process=dict(_COMMON, partial=0.5),
              validity=dict(min_dist_y1=150, max_dist_y1=245, min_dist_y2=140, max_dist_y2=265,
                            min_dist_y3=125, max_dist_y3=290, tangent_thresh=0.46))
SHIPPED_VALIDITY = DEMO_2["validity"]
