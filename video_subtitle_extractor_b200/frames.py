"""Which frames are OCR'd in fast mode, and the half-frame crop as a zero-copy view.

SURVEY.md §8 (f)1 (frame feed), the two pieces that are pure arithmetic:

* `fast_mode_frames`  — `SubtitleExtractor.extract_frame_by_fps`, reference backend/main.py:228-251: one task per
                        `int(fps // extractFrequency)` frames, numbered from 1;
* `half_frame_rows` / `sub_area_view` — `frame_preprocess`, reference backend/tools/subtitle_ocr.py:270-289: the reference
                        slices the lower / upper half of the frame; through the C-ABI that is a pointer offset plus the
                        full frame's row pitch (tests/test_gpu_parity.py::test_sub_area_views_need_no_copy).

Pinned by tests/golden/frames_golden.json (the reference's own functions, tests/golden/make_frames_golden.py).
"""
from __future__ import annotations

from typing import List, Optional, Tuple


def fast_mode_frames(frames_read: int, fps: float, extract_frequency: int = 3) -> List[int]:
    """1-based numbers of the frames that become OCR tasks when the decoder delivers `frames_read` frames."""
    step = max(int(fps // extract_frequency), 1)     # the reference reads step - 1 frames after every task (none if <= 0)
    return list(range(1, frames_read + 1, step))


def half_frame_rows(kind: Optional[str], h: int) -> Tuple[int, int]:
    """Row range [first, end) the reference keeps: 'lower' -> [h // 2, h), 'upper' -> [0, h // 2), anything else -> all."""
    if kind == "lower":
        return h // 2, h
    if kind == "upper":
        return 0, h // 2
    return 0, h


def sub_area_view(ptr: int, h: int, w: int, row_stride: int, rows: Tuple[int, int], cols: Optional[Tuple[int, int]] = None,
                  bytes_per_pixel: int = 3) -> Tuple[int, int, int, int]:
    """(pointer, h, w, row_stride) of a sub-area of a frame for `vse_run` / `Engine.run_device`: no copy, the area keeps the
    frame's row pitch.  Boxes come back in the area's own coordinates, as after the reference's slice."""
    r0, r1 = rows
    c0, c1 = cols if cols is not None else (0, w)
    if not (0 <= r0 < r1 <= h and 0 <= c0 < c1 <= w):
        raise ValueError("empty or out-of-frame sub-area")
    return ptr + r0 * row_stride + c0 * bytes_per_pixel, r1 - r0, c1 - c0, row_stride
