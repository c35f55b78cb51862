#!/usr/bin/env python3
"""Small workload for compute-sanitizer (memcheck / racecheck): two streams, three frames, both filter paths."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lane_tracker_b200 import BatchedLaneTracker, synth  # noqa: E402

cal = synth.shipped_calibration()
t = BatchedLaneTracker(2, **cal)
t.set_capture(True)
vid = synth.RoadVideo(0)
noise = np.random.default_rng(0).integers(0, 256, (720, 1280, 3), dtype=np.uint8)
for i in range(3):
    fr = np.stack([vid.frame(i), noise])            # stream 1 always fails -> attempt 2 path
    d = torch.from_numpy(fr).cuda()
    out = torch.empty_like(d)
    res = t.process(d, out, mask_noise=bool(i == 2))
print("ok", res["valid_lane_lines"], res["attempts"])
t.close()

# two batches in flight (front / back halves on two streams, both intermediate buffer sets)
from lane_tracker_b200 import DevicePipeline, LaneTracker  # noqa: E402
t = BatchedLaneTracker(2, **cal)
pipe = DevicePipeline(t)
batches = [torch.from_numpy(np.stack([vid.frame(i), noise])).cuda() for i in range(4)]
outs = [torch.empty_like(b) for b in batches]
for b, o in zip(batches, outs):
    pipe.submit(b, o)
print("pipelined ok", pipe.fetch_results(2)["valid_lane_lines"])
t.close()

# debug views: sliding-window view, band view, split view with the resize
lt = LaneTracker(**cal)
for i in range(2):
    canvas = lt.process(vid.frame(i), split_view=True)
print("debug views ok", canvas.shape)

# decoder output ingest
nv = torch.from_numpy(np.random.default_rng(1).integers(0, 256, (2, 1080, 1280), dtype=np.uint8)).cuda()
t = BatchedLaneTracker(2, **cal)
print("nv12 ok", tuple(t.nv12_to_rgb(nv).shape))
t.close()
