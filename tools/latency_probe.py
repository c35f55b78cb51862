#!/usr/bin/env python3
"""BASELINE.json configs[1]: one stream, sequential frames, band-search tracking with per-frame state carry.
Reports frames/s and per-frame latency for S in {1, 8, 64} (device-resident frames, overlay rendered)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lane_tracker_b200 import BatchedLaneTracker, synth  # noqa: E402


def run(S, frames=300, pool=4):
    dev = torch.device("cuda", 0)
    host = synth.render_streams(S, pool, workers=8)
    d = torch.from_numpy(host).to(dev).permute(1, 0, 2, 3, 4).contiguous()
    out = torch.empty_like(d[0])
    t = BatchedLaneTracker(S, **synth.shipped_calibration())
    for i in range(5):
        t.process_async(d[i % pool], out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(frames):
        t.process_async(d[i % pool], out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    res = t.fetch_results(S)
    t.close()
    return dict(streams=S, frames_per_stream=frames, ms_per_frame_step=ms / frames, frames_per_s=S * frames / (ms * 1e-3),
                valid=float(res["valid_lane_lines"].mean()), band=float((res["search_mode"] == 1).mean()))


if __name__ == "__main__":
    for S in (1, 8, 64):
        print(json.dumps(run(S, 1000 if S == 1 else 300)))


def run_graph(S, frames=1000, pool=4):
    """Same, with the per-frame chain captured once into a CUDA graph (lane_tracker_b200.GraphedProcess)."""
    from lane_tracker_b200 import GraphedProcess
    dev = torch.device("cuda", 0)
    host = synth.render_streams(S, pool, workers=8)
    d = torch.from_numpy(host).to(dev).permute(1, 0, 2, 3, 4).contiguous()
    t = BatchedLaneTracker(S, **synth.shipped_calibration())
    g = GraphedProcess(t, S)
    for i in range(5):
        g.frames.copy_(d[i % pool], non_blocking=True)
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(frames):
        g.frames.copy_(d[i % pool], non_blocking=True)
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    res = g.fetch_results()
    t.close()
    return dict(streams=S, graph=True, ms_per_frame_step=ms / frames, frames_per_s=S * frames / (ms * 1e-3),
                valid=float(res["valid_lane_lines"].mean()), counter=int(res["counter"][0]))


if __name__ == "__main__":
    for S in (1, 8):
        print(json.dumps(run_graph(S)))
