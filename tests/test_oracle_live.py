"""Oracle vs the LIVE reference (only in the build container, where /root/reference exists)."""
import warnings

import numpy as np
import pytest

import _fixtures as fx
import _liveref
from lane_tracker_b200 import synth
from oracle.tracker import OracleLaneTracker

pytestmark = pytest.mark.skipif(not _liveref.available(), reason="/root/reference not present")


def test_process_matches_live_reference_on_a_sequence():
    warnings.simplefilter("ignore")
    vid = synth.RoadVideo(3)
    ref = _liveref.make_tracker()
    orc = OracleLaneTracker(**synth.shipped_calibration(), backend="cv2")
    frames = [vid.frame(t) for t in range(4)] + [fx.load_frame("test4.jpg")] * 2 + [vid.frame(6)]
    for f in frames:
        a = _liveref.quiet(ref.process, f.copy())
        b = orc.process(f.copy())
        assert fx.out_digest(a) == fx.out_digest(b)
        assert ref.last_detection == orc.last_detection and ref.valid_lane_lines == orc.valid_lane_lines
        for nm in ("left_x", "left_y", "right_x", "right_y", "left_avg_x", "right_avg_x"):
            assert np.array_equal(getattr(ref, nm), getattr(orc, nm)), nm
        assert ref.average_curve_radius == orc.average_curve_radius
        assert ref.eccentricity == orc.eccentricity
