#!/usr/bin/env python3
"""Generate lane_tracker_b200/data/hershey_simplex_s1_t2_aa.npz: per-character pixel sprites of
cv2.putText(FONT_HERSHEY_SIMPLEX, fontScale=1, color=(255,255,255), thickness=2, LINE_AA), the only text style the
reference uses (lane_tracker.py:653-659, 668-672).

OpenCV renders every glyph as anti-aliased thick polylines blended onto the image.  Probing shows (and this script
asserts) that the result is (a) a per-pixel function of the background value only, (b) invariant under integer
translation of the text origin, and (c) for a string, the composition of its characters' functions in order with
integer advances.  So a glyph is fully described by its touched pixels (dy, dx relative to the origin) and, per
pixel, a 256-entry table out = T[background]; across the 95 printable ASCII glyphs only ~1.5k distinct tables occur.

Needs cv2 (the reference's own third-party dependency); the product only reads the generated data file.
"""
import os
import sys

import cv2
import numpy as np

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "lane_tracker_b200", "data",
                   "hershey_simplex_s1_t2_aa.npz")
FONT, SCALE, THICK = cv2.FONT_HERSHEY_SIMPLEX, 1, 2
H, W, OX, OY = 72, 96, 24, 48
FIRST, LAST = 32, 126


def put(img, text, org):
    cv2.putText(img, text, org, FONT, fontScale=SCALE, color=(255, 255, 255), thickness=THICK, lineType=cv2.LINE_AA)
    return img


def glyph_tables(ch):
    L = np.zeros((256, H, W), np.uint8)
    for v in range(256):
        L[v] = put(np.full((H, W, 3), v, np.uint8), ch, (OX, OY))[:, :, 0]
    return L


def advance(ch):
    if ch == " ":
        a = put(np.zeros((H, 3 * W, 3), np.uint8), "| |", (OX, OY))[:, :, 0]
        b = put(np.zeros((H, 3 * W, 3), np.uint8), "||", (OX, OY))[:, :, 0]
        return int(np.nonzero(a.any(0))[0].max() - np.nonzero(b.any(0))[0].max())
    a = put(np.zeros((H, 3 * W, 3), np.uint8), ch + "|", (OX, OY))[:, :, 0]
    b = put(np.zeros((H, 3 * W, 3), np.uint8), "|", (OX, OY))[:, :, 0]
    # position of the trailing bar relative to a bar drawn at the origin = advance of ch
    only_bar = put(np.zeros((H, 3 * W, 3), np.uint8), ch, (OX, OY))[:, :, 0]
    diff = (a.astype(int) != only_bar.astype(int)).any(0)
    return int(np.nonzero(diff)[0].min() - np.nonzero(b.any(0))[0].min())


def main():
    ident = np.arange(256, dtype=np.uint8)[:, None, None]
    tables, index = [], {}
    starts, dys, dxs, lut_idx, adv = [0], [], [], [], []
    for code in range(FIRST, LAST + 1):
        ch = chr(code)
        L = glyph_tables(ch)
        ys, xs = np.nonzero((L != ident).any(0))
        for y, x in zip(ys, xs):
            key = L[:, y, x].tobytes()
            if key not in index:
                index[key] = len(tables)
                tables.append(np.frombuffer(key, np.uint8))
            dys.append(y - OY)
            dxs.append(x - OX)
            lut_idx.append(index[key])
        starts.append(len(dys))
        adv.append(advance(ch))
    np.savez_compressed(OUT, tables=np.stack(tables), char_start=np.array(starts, np.int32),
                        dy=np.array(dys, np.int16), dx=np.array(dxs, np.int16), lut=np.array(lut_idx, np.uint16),
                        advance=np.array(adv, np.int32), first_char=np.int32(FIRST))
    print("wrote", OUT, os.path.getsize(OUT), "bytes;", len(tables), "tables,", len(dys), "pixels")
    # self-check against cv2 on random strings, origins and backgrounds
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from lane_tracker_b200.text import TextSprites
    from oracle.text import render
    sp = TextSprites.load()
    rng = np.random.default_rng(0)
    alphabet = [chr(c) for c in range(FIRST, LAST + 1)]
    for t in range(300):
        n = int(rng.integers(1, 30))
        s = "".join(rng.choice(alphabet, n))
        if t < 4:
            s = ["Curve Radius: 10403 m", "Eccentricity: -0.07 m", "Frame: 970", "Lane Line Detection Failed"][t]
        org = (int(rng.integers(0, 60)), int(rng.integers(30, 200)))
        bg = rng.integers(0, 256, (240, 1280, 3), dtype=np.uint8)
        want = put(bg.copy(), s, org)
        got = render(sp, bg.copy(), s, org)
        assert np.array_equal(want, got), (t, s, org, int((want != got).sum()))
    print("self-check ok: 300 strings bit-exact vs cv2.putText")


if __name__ == "__main__":
    main()
