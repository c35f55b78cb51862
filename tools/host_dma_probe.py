#!/usr/bin/env python3
"""What bounds the host-to-host (`e2e`) throughput?  N ranks (one per GPU, each pinned to its own slice of the host
cores) move the bytes the benchmark's HostPipeline moves -- nothing else -- and the aggregate is compared with the
benchmark's e2e figure:

    python tools/host_dma_probe.py --gpus N > profiles/r02_host_dma_N.jsonl

Per rank and in aggregate, GB/s of
  h2d / d2h            64 whole 1280x720x3 frames from / to pinned host memory (cudaMemcpyAsync)
  h2d+d2h              both directions at once on two streams (what a full-frame pipeline needs every batch)
  roi h2d / d2h / both the row-ROI form of HostPipeline(inplace=True): one cudaMemcpy2DAsync per direction over the
                       frame rows the tracker reads (238 of 720) / the overlay can change (237 of 720)
  host memcpy          a plain CPU copy between two pinned buffers on the rank's cores (host DRAM share per rank)
All ranks start every measurement together (barrier) and run it for the same number of repetitions.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

S, H, W = 64, 720, 1280
FRAME = H * W * 3


def worker(rank, world, barrier, queue, reps):
    try:
        cores = sorted(os.sched_getaffinity(0))
        per = max(1, len(cores) // world)
        mine = cores[rank * per:(rank + 1) * per] or cores
        os.sched_setaffinity(0, set(mine))
    except Exception:
        mine = []
    import numpy as np
    import torch
    from lane_tracker_b200 import BatchedLaneTracker, synth
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    h_in = torch.empty((S, H, W, 3), dtype=torch.uint8).pin_memory()
    h_out = torch.empty((S, H, W, 3), dtype=torch.uint8).pin_memory()
    h_in.fill_(7)
    d_in = torch.empty((S, H, W, 3), dtype=torch.uint8, device=dev)
    d_out = torch.empty((S, H, W, 3), dtype=torch.uint8, device=dev)
    trk = BatchedLaneTracker(S, **synth.shipped_calibration(), device=rank)
    g = trk.geometry
    rows_in = (max(0, min(g["source_rows"][0], g["overlay_rows"][0]) - 2), min(H, max(g["source_rows"][1], g["overlay_rows"][1]) + 2))
    rows_out = g["overlay_rows"]
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def run(name, fn_a, fn_b, bytes_a, bytes_b):
        for _ in range(3):
            if fn_a:
                with torch.cuda.stream(s1):
                    fn_a()
            if fn_b:
                with torch.cuda.stream(s2):
                    fn_b()
        torch.cuda.synchronize(dev)
        barrier.wait()
        t0 = time.perf_counter()
        for _ in range(reps):
            if fn_a:
                with torch.cuda.stream(s1):
                    fn_a()
            if fn_b:
                with torch.cuda.stream(s2):
                    fn_b()
        torch.cuda.synchronize(dev)
        dt = time.perf_counter() - t0
        barrier.wait()
        queue.put((name, rank, (bytes_a + bytes_b) * reps / dt / 1e9, bytes_a * reps / dt / 1e9, bytes_b * reps / dt / 1e9))

    up = lambda: d_in.copy_(h_in, non_blocking=True)
    down = lambda: h_out.copy_(d_out, non_blocking=True)
    rup = lambda: trk.copy_rows(d_in, h_in, rows_in[0], rows_in[1], True)
    rdown = lambda: trk.copy_rows(h_out, d_out, rows_out[0], rows_out[1], False)
    nb_in = S * (rows_in[1] - rows_in[0]) * W * 3
    nb_out = S * (rows_out[1] - rows_out[0]) * W * 3
    run("h2d", up, None, S * FRAME, 0)
    run("d2h", None, down, 0, S * FRAME)
    run("h2d+d2h", up, down, S * FRAME, S * FRAME)
    run("roi_h2d", rup, None, nb_in, 0)
    run("roi_d2h", None, rdown, 0, nb_out)
    run("roi_h2d+d2h", rup, rdown, nb_in, nb_out)
    # host DRAM share: CPU copy between the two pinned buffers on this rank's cores
    a, b = h_in.numpy().reshape(-1), h_out.numpy().reshape(-1)
    barrier.wait()
    t0 = time.perf_counter()
    for _ in range(max(1, reps // 4)):
        np.copyto(b, a)
    dt = time.perf_counter() - t0
    barrier.wait()
    queue.put(("host_memcpy", rank, 2 * a.nbytes * max(1, reps // 4) / dt / 1e9, 0.0, 0.0))
    queue.put(("meta", rank, float(len(mine)), float(rows_in[1] - rows_in[0]), float(rows_out[1] - rows_out[0])))
    trk.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--reps", type=int, default=40)
    args = ap.parse_args()
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    barrier = ctx.Barrier(args.gpus)
    queue = ctx.Queue()
    procs = [ctx.Process(target=worker, args=(r, args.gpus, barrier, queue, args.reps)) for r in range(args.gpus)]
    for p in procs:
        p.start()
    rows = []
    for _ in range(args.gpus * 8):
        rows.append(queue.get(timeout=600))
    for p in procs:
        p.join()
    names = ["h2d", "d2h", "h2d+d2h", "roi_h2d", "roi_d2h", "roi_h2d+d2h", "host_memcpy"]
    meta = [r for r in rows if r[0] == "meta"]
    print(json.dumps({"gpus": args.gpus, "host_cores": os.cpu_count(), "cores_per_rank": meta[0][2], "roi_rows_in": meta[0][3],
                      "roi_rows_out": meta[0][4], "batch": "%d frames of %dx%dx3" % (S, W, H), "reps": args.reps}))
    for n in names:
        r = sorted([x for x in rows if x[0] == n], key=lambda x: x[1])
        print(json.dumps({"test": n, "aggregate_GBps": round(sum(x[2] for x in r), 2), "aggregate_up_GBps": round(sum(x[3] for x in r), 2),
                          "aggregate_down_GBps": round(sum(x[4] for x in r), 2), "per_rank_GBps": [round(x[2], 2) for x in r]}))


if __name__ == "__main__":
    main()
