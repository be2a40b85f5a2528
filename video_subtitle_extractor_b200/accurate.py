"""Accurate-mode frame selection: which frames of a video become OCR tasks.

SURVEY.md §8 (f)2.  The reference walks the video once, runs the DETECTOR on every frame and the full OCR only where it
needs text: at the first frame of a subtitle and on every following frame until the text differs from that first frame
(`SubtitleExtractor.extract_frame_by_det`, reference backend/main.py:255-376, with `_compare_ocr_result` :924-952 and
`__get_area_text` :905-922).  It queues one task for the first and one for the last frame of every subtitle.

With the B200 engine both networks are cheap enough to run over the whole video in large batches (`vse_run`), so the
controller becomes a REPLAY: `detect(k)` / `predict(k)` below read the precomputed results of frame k, and this module
reproduces the reference's decisions — including its quirks: the frame whose text differs ends the previous subtitle but
does not open the next one; a queued task carries the cached OCR result of the frame the loop is at when the task is
flushed (not of the task's own frame), or nothing; without a subtitle area no subtitle is ever opened.

Pinned by tests/golden/accurate_golden.json (the reference's own methods driven by scripted per-frame results,
tests/golden/make_accurate_golden.py).
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

from .dedup import ratio
from .rawtxt import Coordinate, get_coordinates

Task = Tuple[int, Optional[list], Optional[list]]        # (frame number, dt_box | None, rec_res | None)


def _inside(area: Coordinate, c: Coordinate) -> bool:
    """Box (xmin, xmax, ymin, ymax) entirely inside the subtitle area (same layout)."""
    return area[0] <= c[0] and c[1] <= area[1] and area[2] <= c[2] and c[3] <= area[3]


def area_text(dt_box, rec_res, sub_area: Optional[Coordinate]) -> str:
    """Concatenated texts of the boxes inside the subtitle area ('' without an area, as in the reference)."""
    if sub_area is None:
        return ""
    return "".join(t[0] for t, c in zip(rec_res, get_coordinates(dt_box)) if _inside(sub_area, c))


def accurate_mode_tasks(frame_count: int, detect: Callable[[int], Sequence], predict: Callable[[int], Tuple[list, list]],
                        sub_area: Optional[Coordinate], threshold: float = 0.8, frames_read: Optional[int] = None) -> List[Task]:
    """Frames are numbered from 1.  `detect(k)` -> the detector's quads of frame k (anything `list()` turns into a list of
    [4][2] corner lists; empty when nothing was found); `predict(k)` -> (dt_box, rec_res) as `OcrRecogniser.predict` returns
    them (rawtxt.order_like_predict on the engine's result).  `frame_count` is the container's frame count (decides the
    "last frame" rule), `frames_read` how many frames the decoder really delivers (default: the same)."""
    tasks: List[Task] = []
    cache = {}                    # frame number -> (text inside the area, dt_box, rec_res)
    pending: List[int] = []       # frame numbers waiting to be queued
    first, seek_start, seek_end, start_no, cur = True, False, False, 0, 0

    def ocr(k: int) -> str:
        if k not in cache:
            dt_box, rec = predict(k)
            cache[k] = (area_text(dt_box, rec, sub_area), dt_box, rec)
        return cache[k][0]

    def flush(keep: int) -> None:
        while len(pending) > keep:
            no = pending.pop(0)
            hit = cache.get(cur)
            tasks.append((no, hit[1], hit[2]) if hit is not None else (no, None, None))

    for cur in range(1, (frame_count if frames_read is None else frames_read) + 1):
        quads = detect(cur)
        quads = quads.tolist() if hasattr(quads, "tolist") else list(quads)
        if sub_area is not None:
            has_text = any(_inside(sub_area, c) for c in get_coordinates(quads))
            if has_text and first:
                seek_start, first = True, False
        else:
            has_text = len(quads) > 0
        if has_text:
            if seek_start:
                start_no = cur
                fresh = cur not in cache
                ocr(cur)
                if fresh:
                    pending.append(cur)
                seek_start, seek_end = False, True
            if seek_end and cur == frame_count:
                seek_end = False
                pending.append(cur)
            if seek_end:
                same = ratio(ocr(start_no), ocr(cur)) > threshold
                for k in [k for k in cache if k < min(start_no, cur) - 10]:
                    del cache[k]
                if not same:                       # the previous frame was the subtitle's last one
                    seek_end, seek_start = False, True
                    pending.append(cur - 1)
        elif seek_end:
            seek_end, seek_start = False, True
            pending.append(cur - 1)
        flush(keep=1)
    flush(keep=0)
    return tasks
