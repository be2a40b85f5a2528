"""raw.txt post-processing that follows the OCR pass: merge the lines of one frame, collapse runs of similar lines into
(start frame, end frame, text) subtitles.

SURVEY.md §8 (f)3, restated from the reference so that a maintainer who produces raw.txt lines from batched engine
results (rawtxt.py) can finish the job without the reference's file round trip:

* `concat_same_frame`  — `SubtitleExtractor._concat_content_with_same_frameno`, reference backend/main.py:820-864
* `remove_duplicates`  — `SubtitleExtractor._remove_duplicate_subtitle`,        reference backend/main.py:774-818
* `ratio`              — `Levenshtein.ratio` (third-party, absent here): normalised indel similarity
                         2 LCS(a, b) / (len(a) + len(b)), 1.0 for two empty strings

Both work on lists of raw.txt lines (``"{frame:08d}\\t{coordinate}\\t{text}\\n"``) instead of rewriting the file in place.
Pinned by tests/golden/dedup_golden.json: outputs of the reference's own methods on seeded files
(tests/golden/make_dedup_golden.py; the similarity function is pinned by its definition only).
"""
from __future__ import annotations

import unicodedata
from typing import List, Sequence, Tuple


def ratio(a: str, b: str) -> float:
    if not a and not b:
        return 1.0
    if len(a) < len(b):
        a, b = b, a
    row = [0] * (len(b) + 1)             # LCS by rows, O(len(a) * len(b)) time, O(len(b)) space
    for x in a:
        diag = 0
        for j, y in enumerate(b, 1):
            diag, row[j] = row[j], (diag + 1 if x == y else max(row[j], row[j - 1]))
    return 2.0 * row[-1] / (len(a) + len(b))


def _fields(line: str) -> List[str]:
    p = line.split("\t")
    return [p[0], p[1], p[2]]


def concat_same_frame(lines: Sequence[str]) -> List[str]:
    """Frames with several text lines become ONE line: the texts joined by a blank in file order (every text keeps the blank
    its newline turns into, exactly as the reference writes it), carried by the first line's coordinate; every text is
    NFKC-normalised on the way out."""
    rows = [_fields(l) for l in lines]
    seen, multi = {}, []
    for r in rows:
        seen[r[0]] = seen.get(r[0], 0) + 1
    for r in rows:
        if seen[r[0]] > 1 and r[0] not in multi:
            multi.append(r[0])
    doomed = []
    for frame in multi:
        members = [r for r in rows if r[0] == frame]
        joined = " ".join(m[2] for m in members).replace("\n", " ") + "\n"
        for m in members:
            m[2] = joined
        doomed += members[1:]
    for d in doomed:                      # by VALUE, first match, as list.remove does in the reference
        if d in rows:
            rows.remove(d)
    return [f"{f}\t{c}\t{unicodedata.normalize('NFKC', t)}" for f, c, t in rows]


def remove_duplicates(lines: Sequence[str], threshold: float = 0.8, use_vsf: bool = False) -> List[Tuple[str, str, str]]:
    """-> [(start frame, end frame, text)].  A subtitle is a maximal run of consecutive lines whose blank-stripped text stays
    similar (ratio >= threshold) to the run's FIRST line; its text is the longest (blank-stripped) line of the run, first
    one on ties.  Without VideoSubFinder a one-line run ends at the next line's frame (unless it is the last line)."""
    rows = [(f, t) for f, _, t in map(_fields, concat_same_frame(lines))]
    out, i, n = [], 0, len(rows)
    while i < n:
        head = rows[i][1].replace(" ", "")
        j = i
        while j + 1 < n and not ratio(head, rows[j + 1][1].replace(" ", "")) < threshold:
            j += 1
        start, end = rows[i][0], rows[j][0]
        if not use_vsf and end == start and j + 1 < n:
            end = rows[j + 1][0]
        run = rows[i:j + 1]
        best = max(range(len(run)), key=lambda k: len(run[k][1].replace(" ", "")))
        out.append((start, end, run[best][1]))
        i = j + 1
    return out


# ------------------------------------------------------------------------------------------------------------------------
# SRT writer — `SubtitleExtractor.generate_subtitle_file` / `_frame_to_timecode`, reference backend/main.py:614-636, 731-766
# (the VideoSubFinder variant of the writer is out of scope with VideoSubFinder itself).  Pinned by
# tests/golden/srt_golden.json: the reference's own methods on its sample video test/test_en.mp4.
# ------------------------------------------------------------------------------------------------------------------------
def timecode_from_frame(frame_no: int, fps: float) -> str:
    """Fallback of the reference when the decoder cannot deliver the frame: HH:MM:SS from the frame count and, as the last
    field, the frame's index inside its second (`frame_no % fps`) — not milliseconds; kept as the reference writes it."""
    return "{0:02d}:{1:02d}:{2:02d},{3:03d}".format(int(frame_no / (3600 * fps)), int(frame_no / (60 * fps) % 60),
                                                    int(frame_no / fps % 60), int(frame_no % fps))


def timecode_from_msec(msec: float) -> str:
    """HH:MM:SS,mmm from the decoder's position in milliseconds (truncated, as the reference does)."""
    seconds, ms = msec // 1000, int(msec % 1000)
    minutes = hours = 0
    if seconds >= 60:
        minutes, seconds = int(seconds // 60), int(seconds % 60)
    if minutes >= 60:
        hours, minutes = int(minutes // 60), int(minutes % 60)
    return "%02d:%02d:%02d,%03d" % (hours, minutes, seconds, ms)


def srt_text(subtitles: Sequence[Tuple[str, str, str]], fps: float, pos_msec=None) -> Tuple[str, List[int]]:
    """-> (contents of the .srt file, 1-based numbers of the subtitles that were stretched to one second).

    `subtitles`: output of `remove_duplicates`.  `pos_msec(frame_no)` returns the decoder's timestamp of that frame in
    milliseconds (what cv2 reports as CAP_PROP_POS_MSEC after seeking to the frame and reading it) or None when the frame
    cannot be read; without it, or for a timestamp <= 0, the frame-count fallback is used.  A subtitle shorter than one
    second of frames ends one second (`int(start + fps)` frames) after its start."""
    def timecode(frame_no: int) -> str:
        ms = pos_msec(frame_no) if pos_msec is not None else None
        return timecode_from_frame(frame_no, fps) if ms is None or ms <= 0 else timecode_from_msec(ms)

    out, stretched = [], []
    for number, (start, end, text) in enumerate(subtitles, 1):
        first, last = int(start), int(end)
        if abs(last - first) < fps:
            last = int(first + fps)
            stretched.append(number)
        out.append(f"{number}\n{timecode(first)} --> {timecode(last)}\n{text}\n")
    return "".join(out), stretched
