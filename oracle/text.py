"""CPU statement of the text overlays (test infrastructure, like the rest of oracle/): the composition rule of the glyph
sprites in NumPy -- the checker of the device renderer k_text -- and the strings the reference formats
(lane_tracker.py:653-659, 668-672).  Pinned against cv2.putText by tests/test_oracle_cvops.py."""
import numpy as np


def render(sp, img, text, org):
    """In-place equivalent of cv2.putText(img, text, org, HERSHEY_SIMPLEX, 1, (255,255,255), 2, LINE_AA) from the glyph
    sprites ``sp`` (lane_tracker_b200.text.TextSprites)."""
    h, w = img.shape[:2]
    x0, y0 = org
    for ch in text:
        c = ord(ch) - sp.first_char
        if not (0 <= c < len(sp.advance)):
            c = ord("?") - sp.first_char
        a, b = sp.char_start[c], sp.char_start[c + 1]
        ys = y0 + sp.dy[a:b].astype(np.int64)
        xs = x0 + sp.dx[a:b].astype(np.int64)
        ok = (ys >= 0) & (ys < h) & (xs >= 0) & (xs < w)
        ys, xs, li = ys[ok], xs[ok], sp.lut[a:b][ok].astype(np.int64)
        for chn in range(img.shape[2]):
            img[ys, xs, chn] = sp.tables[li, img[ys, xs, chn]]
        x0 += int(sp.advance[c])
    return img


def overlay_strings(drew_lane, average_curve_radius, eccentricity, counter, print_frame_count):
    """The (text, origin) pairs process() draws on a frame (draw_lane :653-659 / print_failure :668-672)."""
    if drew_lane:
        out = [("Curve Radius: {} m".format(average_curve_radius), (20, 35)),
               ("Eccentricity: {:.2f} m".format(eccentricity), (20, 70))]
        if print_frame_count:
            out.append(("Frame: {}".format(counter - 1), (20, 105)))
    else:
        out = [("Lane Line Detection Failed", (20, 35))]
        if print_frame_count:
            out.append(("Frame: {}".format(counter - 1), (20, 70)))
    return out
