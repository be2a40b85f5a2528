"""Fast-mode frame schedule and half-frame crop (video_subtitle_extractor_b200/frames.py) against the reference's own
`extract_frame_by_fps` / `frame_preprocess` (tests/golden/frames_golden.json, tests/golden/make_frames_golden.py)."""
import json
import os

import numpy as np
import pytest

from video_subtitle_extractor_b200 import frames

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "frames_golden.json")


def test_fast_mode_schedule_and_half_frame_rows_match_reference():
    with open(GOLDEN) as f:
        g = json.load(f)
    assert len(g["schedule"]) >= 100
    for s in g["schedule"]:
        assert frames.fast_mode_frames(s["n_frames"], s["fps"], s["extract_frequency"]) == s["frames"], s
    for c in g["crops"]:
        r0, r1 = frames.half_frame_rows(c["kind"], c["h"])
        assert r1 - r0 == c["n_rows"] and (c["first_row"] is None or c["first_row"] == r0), c


def test_sub_area_view_addresses_the_same_pixels_as_a_numpy_slice():
    frame = np.arange(20 * 30 * 3, dtype=np.uint8).reshape(20, 30, 3)
    ptr, h, w, stride = frames.sub_area_view(frame.ctypes.data, 20, 30, frame.strides[0], frames.half_frame_rows("lower", 20), (4, 29))
    view = frame[10:, 4:29]
    assert (ptr, h, w, stride) == (view.ctypes.data, view.shape[0], view.shape[1], view.strides[0])
    with pytest.raises(ValueError):
        frames.sub_area_view(frame.ctypes.data, 20, 30, frame.strides[0], (10, 10))
