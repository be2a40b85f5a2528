"""Host-side consumer of the engine's per-frame results: reading order, ROI / score filter and the raw.txt wire format.

SURVEY.md §8 rows a1 / a4 (reference-side glue around the predictor call) and (f)3 (first widening row).  The reference
does this per frame in Python right after `predict`; a maintainer who batches frames through `vse_run`
needs the same three steps on the engine's `FrameResult`s, so they are restated here on plain rectangles (the reference
goes through shapely polygons; every polygon it builds is an axis-aligned rectangle):

* `order_like_predict`  — `OcrRecogniser.predict`, reference backend/tools/ocr.py:24-86
* `get_coordinates`     — reference backend/tools/ocr.py:115-134
* `frame_lines`         — `extract_subtitles`, reference backend/tools/subtitle_ocr.py:20-83 (the lines it appends to
                          `raw_subtitles`; the debug dump and console log are not part of the wire format)

Pinned by tests/golden/rawtxt_golden.json: outputs of the reference's own functions on seeded inputs
(tests/golden/make_rawtxt_golden.py).
"""
from __future__ import annotations

import re
from typing import List, Optional, Sequence, Tuple

Coordinate = Tuple[int, int, int, int]          # (xmin, xmax, ymin, ymax)
_CJK_UNIFIED = re.compile("[一-龥]")     # what the reference strips from English subtitles (subtitle_ocr.py:36)


def y_round(y: int) -> int:
    """Nearest multiple of 10, ties and exact multiples going down (ocr.py:16-22)."""
    up, down = y + 10 - y % 10, y - y % 10
    return up if abs(y - up) < abs(y - down) else down


def _inner_rect(quad) -> List[int]:
    """Largest axis-aligned rectangle spanned by the quad's corner ORDER (tl, tr, br, bl), ints by truncation (ocr.py:31-41)."""
    (x1, y1), (x2, y2), (x3, y3), (x4, y4) = [(int(p[0]), int(p[1])) for p in list(quad)[:4]]
    return [max(x1, x4), min(x2, x3), max(y1, y2), min(y3, y4)]


def get_coordinates(dt_box) -> List[Coordinate]:
    """(xmin, xmax, ymin, ymax) per box; anything that is not a list yields [] as in the reference (ocr.py:122)."""
    if not isinstance(dt_box, list):
        return []
    return [tuple(_inner_rect(q)) for q in dt_box]


def order_like_predict(quads, rec_res):
    """-> (dt_box, res) exactly as `OcrRecogniser.predict` hands them to `extract_subtitles`.

    Boxes are bucketed into text lines by their top edge rounded to 10 px (a new line only opens when neither it nor a
    neighbour 10 px away exists), every box's ymin is REPLACED by its line's y, lines go top to bottom and boxes inside a
    line left to right (stable).  dt_box rows are the four corners of the (xmin, xmax, line y, ymax) rectangle."""
    if len(quads) == 0:
        return quads, rec_res
    rects = [_inner_rect(q) for q in quads] if isinstance(quads, list) else []
    lines: List[int] = []
    for r in rects:
        yr = y_round(r[2])
        if not lines or (yr not in lines and yr + 10 not in lines and yr - 10 not in lines):
            lines.append(yr)
    lines.sort()
    for r in rects:
        for ly in lines:                       # re-evaluated after every assignment, like the reference's loop
            if abs(ly - y_round(r[2])) <= 10:
                r[2] = ly
    ranked = []
    for ly in lines:
        members = [(r, t) for r, t in zip(rects, rec_res) if r[2] == ly]
        members.sort(key=lambda m: m[0][0])    # the reference's bubble sort on xmin is a stable ascending sort
        ranked += members
    dt_box = [[(r[0], r[2]), (r[1], r[2]), (r[1], r[3]), (r[0], r[3])] for r, _ in ranked]
    return dt_box, [t for _, t in ranked]


def overflow_rate(sub_area: Coordinate, box: Coordinate) -> Optional[float]:
    """union(sub_area, box) / area(sub_area) - 1, or None when the rectangles do not even touch (subtitle_ocr.py:52-58).
    Both rectangles are (xmin, xmax, ymin, ymax); a box whose min exceeds its max is the same rectangle mirrored."""
    sx0, sx1, sy0, sy1 = min(sub_area[0], sub_area[1]), max(sub_area[0], sub_area[1]), min(sub_area[2], sub_area[3]), max(sub_area[2], sub_area[3])
    bx0, bx1, by0, by1 = min(box[0], box[1]), max(box[0], box[1]), min(box[2], box[3]), max(box[2], box[3])
    ix0, ix1, iy0, iy1 = max(sx0, bx0), min(sx1, bx1), max(sy0, by0), min(sy1, by1)
    if ix0 > ix1 or iy0 > iy1:
        return None
    a_sub, a_box, a_int = float((sx1 - sx0) * (sy1 - sy0)), float((bx1 - bx0) * (by1 - by0)), float((ix1 - ix0) * (iy1 - iy0))
    return ((a_sub + a_box - a_int) / a_sub) - 1


def frame_lines(frame_no: int, dt_box, rec_res: Sequence[Tuple[str, float]], sub_area: Optional[Coordinate] = None,
                rec_char_type: str = "en", drop_score: float = 0.75, sub_area_deviation_rate: float = 0.0) -> List[str]:
    """raw.txt lines of one frame: ``"{frame_no:08d}\\t(xmin, xmax, ymin, ymax)\\t{text}\\n"``.

    With a subtitle area, a line is kept when its rectangle touches the area, sticks out of it by at most
    `sub_area_deviation_rate` (relative to the area's size) and was recognised with a probability ABOVE `drop_score`;
    without one every line is kept (the reference does not even apply the score there).  `sub_area` is
    (xmin, xmax, ymin, ymax) in the same pixel coordinates as the boxes."""
    out = []
    for (text, prob), coord in zip(rec_res, get_coordinates(dt_box)):
        if rec_char_type == "en":
            text = _CJK_UNIFIED.sub("", text)
        if sub_area is not None:
            rate = overflow_rate(sub_area, coord)
            if rate is None or not (rate <= sub_area_deviation_rate and prob > drop_score):
                continue
        out.append(f"{str(frame_no).zfill(8)}\t{coord}\t{text}\n")
    return out


def lines_from_frame_result(frame_no: int, result, characters: Sequence[str], **filter_kw) -> List[str]:
    """raw.txt lines of one `engine.FrameResult` (quads float32 [n,4,2] in TextSystem order, CTC class ids, scores):
    ids -> text with the language's character list (charset.py), then the reference's predict ordering and filter."""
    from .charset import ids_to_text
    quads = [q for q in result.quads]                       # list of [4,2] arrays: what PaddleOCR returns (ocr.py:27-30)
    rec = [(ids_to_text(ids, characters), float(s)) for ids, s in zip(result.ids, result.rec_scores)]
    dt_box, res = order_like_predict(quads, rec)
    return frame_lines(frame_no, dt_box, res, **filter_kw)
