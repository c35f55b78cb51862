#!/usr/bin/env python3
"""Sequential process_async vs DevicePipeline (front half of batch k+1 under the back half of batch k)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lane_tracker_b200 import BatchedLaneTracker, DevicePipeline, synth  # noqa: E402


def main():
    S, P, K = int(os.environ.get("S", 64)), 4, 100
    dev = torch.device("cuda", 0)
    pool = torch.from_numpy(synth.render_streams(S, P, first_seed=0, workers=16)).to(dev).permute(1, 0, 2, 3, 4).contiguous()
    outs = [torch.empty_like(pool[0]) for _ in range(2)]
    res = {}
    for mode in ("sequential", "pipelined"):
        trk = BatchedLaneTracker(S, **synth.shipped_calibration(), device=0)
        pipe = DevicePipeline(trk) if mode == "pipelined" else None

        def step(i):
            if pipe is None:
                trk.process_async(pool[i % P], outs[i & 1])
            else:
                pipe.submit(pool[i % P], outs[i & 1])
        for i in range(4):
            step(i)
        if pipe:
            pipe.join()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(4, 4 + K):
            step(i)
        if pipe:
            pipe.join()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / K
        r = pipe.fetch_results(S) if pipe else trk.fetch_results(S)
        res[mode] = (r, outs[(4 + K - 1) & 1].clone())
        print(mode, "ms/step %.4f  frames/s %.0f  valid %.2f" % (ms, S / ms * 1e3, r["valid_lane_lines"].mean()), flush=True)
        trk.close()
    a, b = res["sequential"], res["pipelined"]
    same = all(np.array_equal(a[0][f], b[0][f]) for f in a[0].dtype.names)
    print("results identical:", same, " frames identical:", bool(torch.equal(a[1], b[1])))


if __name__ == "__main__":
    main()
