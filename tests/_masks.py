"""Random binary masks that stress the sliding-window state machine."""
import numpy as np


def random_masks(n, seed=0, shape=(1100, 1080)):
    rng = np.random.default_rng(seed)
    H, W = shape
    out = []
    for t in range(n):
        m = np.zeros(shape, np.uint8)
        kind = t % 6
        ys = np.arange(H)
        if kind == 0:      # two steep lines drifting towards / beyond the image edges
            for c0, sl in ((rng.uniform(380, 520), rng.uniform(-0.6, -0.1)), (rng.uniform(560, 700), rng.uniform(0.1, 0.6))):
                xs = (c0 + sl * (H - ys)).astype(int)
                for dx in range(-3, 4):
                    ok = (xs + dx >= 0) & (xs + dx < W)
                    m[ys[ok], xs[ok] + dx] = 255
        elif kind == 1:    # dashed right line only in some levels, solid left: exercises the coupling fallback
            xl = (450 + 40 * np.sin(ys / 150.0)).astype(int)
            xr = xl + rng.integers(150, 220)
            on = ((ys // rng.integers(30, 120)) % 2) == 0
            for dx in range(-4, 5):
                m[ys, np.clip(xl + dx, 0, W - 1)] = 255
                m[ys[on], np.clip(xr[on] + dx, 0, W - 1)] = 255
        elif kind == 2:    # sparse salt noise
            m[rng.random(shape) < rng.uniform(0.001, 0.02)] = 255
        elif kind == 3:    # blobs
            for _ in range(rng.integers(5, 40)):
                y, x = rng.integers(0, H), rng.integers(0, W)
                m[max(0, y - 20):y + 20, max(0, x - 6):x + 6] = 255
        elif kind == 4:    # only one side populated
            xs = (rng.uniform(380, 520) + 0.05 * (H - ys)).astype(int)
            m[ys, np.clip(xs, 0, W - 1)] = 255
        else:              # lines that leave the frame on the left quickly (negative window starts)
            xs = (400 - 1.2 * (H - ys)).astype(int)
            ok = (xs >= 0) & (xs < W)
            m[ys[ok], xs[ok]] = 255
            xs2 = (600 - 1.0 * (H - ys)).astype(int)
            ok = (xs2 >= 0) & (xs2 < W)
            m[ys[ok], xs2[ok]] = 255
        out.append(m)
    return out
