#!/usr/bin/env python3
"""Sweep the band split of the paired morphology launch (LT_MORPH_BANDS knob) and print the stage times."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lane_tracker_b200 import BatchedLaneTracker, synth  # noqa: E402


def main():
    S, P = 64, 2
    dev = torch.device("cuda", 0)
    pool = torch.from_numpy(synth.render_streams(S, P, first_seed=0, workers=16)).to(dev).permute(1, 0, 2, 3, 4).contiguous()
    out = torch.empty_like(pool[0])
    combos = [None] + [(a, b) for a in (1, 2, 3) for b in (3, 4, 5, 6, 7, 8, 10)]
    for c in combos:
        if c is None:
            os.environ.pop("LT_MORPH_BANDS", None)
        else:
            os.environ["LT_MORPH_BANDS"] = "%d,%d" % c
        trk = BatchedLaneTracker(S, **synth.shipped_calibration(), device=0)
        for i in range(3):
            trk.process_async(pool[i % P], out)
        torch.cuda.synchronize()
        trk.profile_begin(20)
        for i in range(20):
            trk.process_async(pool[i % P], out)
        torch.cuda.synchronize()
        st, _ = trk.profile_read()
        print(c, "erode %.4f tophat %.4f" % (st["erode55"], st["tophat55"]), flush=True)
        del trk


if __name__ == "__main__":
    main()
