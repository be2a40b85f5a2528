"""Whole-video jobs on the engine: frame feed -> vse_run batches -> raw.txt lines -> de-dup -> .srt text.

SURVEY.md §8 (f): the callers either side of the predictor path, wired to the CUDA engine (the per-frame arithmetic stays in
csrc/; this module only moves frames in and result records out):

* `FrameFeed` — (f)1: the reference decodes one frame per OCR task (`ocr_task_producer`, reference
  backend/tools/subtitle_ocr.py:164-208: seek + read) or walks the video with `cap.read()` (`extract_frame_by_fps` /
  `extract_frame_by_det`, backend/main.py:228-251, 275-283).  Here up to four decoder threads (FrameFeed: contiguous segments of whole batches, one seeking capture each) read the rank's frame range
  sequentially into a ring of pinned batches; the consumer issues `vse_prefetch` for batch k+1 before `vse_run` of batch k,
  so decode, host->device copy and kernels overlap.  The half-frame crop of `frame_preprocess` (subtitle_ocr.py:270-289) is a
  zero-copy view (frames.sub_area_view).
* `fast_mode_job` — BASELINE configs[0]/[3]: `extract_frame_by_fps` schedule -> `vse_run` -> rawtxt.lines_from_frame_result
  (`OcrRecogniser.predict` order + `extract_subtitles` filter) -> shard.gather_by_frame -> dedup.remove_duplicates -> srt_text.
* `accurate_mode_job` — BASELINE configs[2]/[4]: det + rec over EVERY frame in large batches, then the reference's
  `extract_frame_by_det` decisions as a replay (accurate.accurate_mode_tasks) -> the same raw.txt / de-dup / .srt tail.

Frames shard by contiguous ranges over the ranks (shard.frame_range); the only exchange is the gather of result lines.
"""
from __future__ import annotations

import queue
import threading
from dataclasses import dataclass, field
from typing import Callable, Dict, Iterator, List, Optional, Sequence, Tuple

import numpy as np

from . import accurate, dedup, frames as F, rawtxt, shard
from .rawtxt import Coordinate


def default_decoders(world: int = 1) -> int:
    """Decoder threads per rank: up to four, leaving ~four host cores per decoder (every cv2 / ffmpeg capture spawns its own
    frame threads) — fewer when several ranks share the host.  VSE_DECODERS overrides."""
    import os
    if os.environ.get("VSE_DECODERS"):
        return max(1, int(os.environ["VSE_DECODERS"]))
    return max(1, min(4, (os.cpu_count() or 1) // (4 * max(world, 1))))


def default_sub_area(h: int, w: int) -> Coordinate:
    """(xmin, xmax, ymin, ymax) of the reference's default subtitle area: fractions 0.78 / 0.99 / 0.05 / 0.95 (config.py:49)."""
    return int(w * 0.05), int(w * 0.95), int(h * 0.78), int(h * 0.99)


@dataclass
class Batch:
    numbers: List[int]            # 1-based frame numbers, as the reference counts them
    slot: int
    array: np.ndarray             # [n, H, W, 3] uint8 view of the pinned slot
    ptrs: List[int] = field(default_factory=list)


_PINNED_POOL: Dict[int, List[np.ndarray]] = {}   # slot index -> flat page-locked uint8 buffers, reused by successive feeds


def _pinned_slot(index: int, shape) -> np.ndarray:
    """Page-locked [batch, H, W, 3] view for ring slot `index`.  cudaHostAlloc of a 200 MB slot costs ~0.1 s and a job opens one
    feed per video, so the flat buffers are kept and re-viewed (grown when a later video needs more bytes)."""
    import torch                      # container for page-locked host memory only
    need = int(np.prod(shape))
    flat = _PINNED_POOL.get(index)
    if flat is None or flat[0].size < need:
        flat = [torch.empty((need,), dtype=torch.uint8, pin_memory=True).numpy()]
        _PINNED_POOL[index] = flat
    return flat[0][:need].reshape(shape)


class FrameFeed:
    """Batched decode of frames [first, last] (1-based, inclusive) of one video into pinned ring buffers.

    `wanted`: the frame numbers to deliver (None = every frame of the range).  A decoder thread seeks once to the start of its
    range (cv2 CAP_PROP_POS_FRAMES, what the reference's producer does per task) and then only reads forward (`grab()` for the
    frames nobody asked for).  `decoders` > 1 cuts the delivered frames into that many contiguous segments (whole batches) with
    one decoder thread each: one cv2 / ffmpeg capture keeps only a few host cores busy, and with the engine in place decode is
    the limiter of a whole-video job.  Batches then arrive in completion order, not frame order — every consumer here keys
    results by frame number (run_feed)."""

    def __init__(self, path: str, first: int, last: int, wanted: Optional[Sequence[int]] = None, batch: int = 32, slots: int = 3,
                 pinned: bool = True, half: Optional[str] = None, open_capture: Optional[Callable] = None, decoders: int = 1):
        import cv2
        self._open = open_capture or (lambda p: cv2.VideoCapture(p))
        cap = self._open(path)
        if not cap.isOpened():
            raise RuntimeError(f"cannot open {path}")
        self.fps = cap.get(cv2.CAP_PROP_FPS)
        self.frame_count = int(cap.get(cv2.CAP_PROP_FRAME_COUNT))
        self.h, self.w = int(cap.get(cv2.CAP_PROP_FRAME_HEIGHT)), int(cap.get(cv2.CAP_PROP_FRAME_WIDTH))
        cap.release()
        self.path, self.first, self.last, self.batch = path, first, last, batch
        self.wanted = None if wanted is None else sorted(k for k in wanted if first <= k <= last)
        self.rows = F.half_frame_rows(half, self.h)
        # segments: (first frame to decode, last frame to decode, frames to deliver or None = all), whole batches each
        n_deliver = len(self.wanted) if self.wanted is not None else max(0, last - first + 1)
        n_batches = -(-n_deliver // batch) if n_deliver else 0
        k = max(1, min(int(decoders), n_batches)) if n_batches else 1
        self._segments = []
        for j in range(k):
            lo, hi = (j * n_batches // k) * batch, min(((j + 1) * n_batches // k) * batch, n_deliver)
            if self.wanted is not None:
                seg = self.wanted[lo:hi]
                # the first segment starts where the caller said (recorded timestamps cover the skipped lead-in, as before)
                self._segments.append((first if j == 0 else seg[0], seg[-1], seg) if seg else (first, first - 1, []))
            else:
                self._segments.append((first + lo, first + hi - 1, None))
        slots = max(slots, len(self._segments) + 2)
        self.slots = []
        for _ in range(slots):
            if pinned:
                self.slots.append(_pinned_slot(len(self.slots), (batch, self.h, self.w, 3)))
            else:
                self.slots.append(np.empty((batch, self.h, self.w, 3), np.uint8))
        self._free: "queue.Queue[int]" = queue.Queue()
        for s in range(slots):
            self._free.put(s)
        self._ready: "queue.Queue[Optional[Batch]]" = queue.Queue()
        self._read = [0] * len(self._segments)
        self.msec: Dict[int, float] = {}          # decoder timestamp after reading frame `no` (what the SRT writer asks for)
        self._err: Optional[BaseException] = None
        self._threads = [threading.Thread(target=self._decode, args=(j,), daemon=True) for j in range(len(self._segments))]
        for t in self._threads:
            t.start()

    @property
    def frames_read(self) -> int:
        return sum(self._read)

    def _decode(self, j: int):
        import cv2
        seg_first, seg_last, seg_wanted = self._segments[j]
        try:
            cap = self._open(self.path)
            if seg_first > 1:
                cap.set(cv2.CAP_PROP_POS_FRAMES, seg_first - 1)
            want = None if seg_wanted is None else set(seg_wanted)
            stop = seg_last if want is None else (max(want) if want else 0)
            no = seg_first - 1
            cur: Optional[Batch] = None
            msec = {}
            while no < stop:
                if not cap.grab():                 # frames that are not OCR tasks are decoded but never converted / copied
                    break
                no += 1
                self._read[j] += 1
                msec[no] = cap.get(cv2.CAP_PROP_POS_MSEC)
                if want is not None and no not in want:
                    continue
                ok, frame = cap.retrieve()
                if not ok:
                    break
                if cur is None:
                    s = self._free.get()
                    cur = Batch([], s, self.slots[s])
                cur.array[len(cur.numbers)] = frame
                cur.numbers.append(no)
                if len(cur.numbers) == self.batch:
                    self._ready.put(cur)
                    cur = None
            if cur is not None and cur.numbers:
                self._ready.put(cur)
            cap.release()
            self.msec.update(msec)
        except BaseException as e:          # surfaced in the consumer
            self._err = e
        self._ready.put(None)

    def view(self, b: Batch):
        """(ptrs, heights, widths, strides) of the batch's frames — or of their half-frame views — for vse_run / vse_prefetch."""
        r0, r1 = self.rows
        n = len(b.numbers)
        stride = self.w * 3
        base = b.array.ctypes.data
        ptrs = [base + j * self.h * stride + r0 * stride for j in range(n)]
        return ptrs, [r1 - r0] * n, [self.w] * n, [stride] * n

    def release(self, b: Batch):
        self._free.put(b.slot)

    def __iter__(self) -> Iterator[Batch]:
        done = 0
        while done < len(self._threads):
            b = self._ready.get()
            if b is None:
                done += 1
                if self._err is not None:
                    raise self._err
                continue
            yield b


def run_feed(engine, feed: FrameFeed, det_only: bool = False, mem_kind: Optional[int] = None,
             stats: Optional[dict] = None) -> Dict[int, object]:
    """Every delivered frame through vse_run (det_only: vse_det_only) with one batch of look-ahead on the copy stream.
    -> {frame number: engine.FrameResult}.  stats (optional) accumulates seconds spent waiting for the decoder ('feed_wait_s')
    and inside the engine calls ('engine_s') and the frames decoded / delivered."""
    import time
    from . import engine as E
    mk = E.MEM_PINNED if mem_kind is None else mem_kind
    out: Dict[int, object] = {}
    it = iter(feed)
    t_wait = t_eng = 0.0
    t0 = time.perf_counter()
    cur = next(it, None)
    t_wait += time.perf_counter() - t0
    while cur is not None:
        t0 = time.perf_counter()
        nxt = next(it, None)                       # decoded (or being decoded) while the current batch computes
        t1 = time.perf_counter()
        if nxt is not None:
            engine.prefetch(*feed.view(nxt)[:4], mem_kind=mk)
        ptrs, hs, ws, st = feed.view(cur)
        res = engine.run_device(ptrs, hs, ws, st, det_only=det_only, mem_kind=mk)
        t2 = time.perf_counter()
        t_wait += t1 - t0
        t_eng += t2 - t1
        for no, r in zip(cur.numbers, res):
            out[no] = r
        feed.release(cur)
        cur = nxt
    if stats is not None:
        stats["feed_wait_s"] = stats.get("feed_wait_s", 0.0) + t_wait
        stats["engine_s"] = stats.get("engine_s", 0.0) + t_eng
        stats["frames_decoded"] = stats.get("frames_decoded", 0) + feed.frames_read
        stats["frames_ocr"] = stats.get("frames_ocr", 0) + len(out)
    return out


def _srt(lines: List[str], fps: float, path: str, threshold: float,
         msec: Optional[Dict[int, float]] = None) -> Tuple[List[Tuple[str, str, str]], str]:
    """`msec`: timestamps the feed recorded while decoding ({1-based frame number: POS_MSEC after reading it}).  The reference
    seeks and reads once per subtitle boundary (`_frame_to_timecode`, backend/main.py:738-742: POS_FRAMES = n, read, POS_MSEC);
    that is the timestamp after reading frame n + 1, which the sequential decode has already seen — identical values, no seek.
    Frames the feed did not decode fall back to the reference's seek."""
    import cv2
    subs = dedup.remove_duplicates(lines, threshold, use_vsf=False)
    cap = [None]

    def pos_msec(frame_no):
        if msec is not None and frame_no + 1 in msec:
            return msec[frame_no + 1]
        if cap[0] is None:
            cap[0] = cv2.VideoCapture(path)
        cap[0].set(cv2.CAP_PROP_POS_FRAMES, frame_no)
        ok, _ = cap[0].read()
        return cap[0].get(cv2.CAP_PROP_POS_MSEC) if ok else None

    text, _ = dedup.srt_text(subs, fps, pos_msec)
    if cap[0] is not None:
        cap[0].release()
    return subs, text


def _gather_msec(feed: Optional["FrameFeed"]) -> Dict[int, float]:
    local = sorted(feed.msec.items()) if feed is not None else []
    return dict(shard.gather_by_frame(local))


@dataclass
class JobResult:
    lines: List[str]                       # raw.txt lines of the whole video, frame-ordered (every rank holds them)
    subtitles: List[Tuple[str, str, str]]
    srt: str
    frames_ocr: int                        # frames this rank pushed through the engine
    frame_numbers: List[int]               # ... and their numbers
    results: Dict[int, object]


def fast_mode_job(engine, path: str, characters: Sequence[str], rank: int = 0, world: int = 1, batch: int = 32,
                  sub_area: Optional[Coordinate] = "default", rec_char_type: str = "en", drop_score: float = 0.75,
                  extract_frequency: int = 3, threshold: float = 0.8, half: Optional[str] = None, pinned: bool = True,
                  write_srt: bool = True, stats: Optional[dict] = None, decoders: Optional[int] = None) -> JobResult:
    """Fast mode of the reference (`run` -> `extract_frame_by_fps`, backend/main.py:145-147) on this rank's share of the
    schedule.  `sub_area` 'default' = the reference's default area for the video's size; None = no area."""
    import time
    clock = [time.perf_counter()]

    def lap(key):       # seconds since the previous lap, accumulated in stats[key]
        now = time.perf_counter()
        if stats is not None:
            stats[key] = stats.get(key, 0.0) + now - clock[0]
        clock[0] = now

    probe = FrameFeed(path, 1, 0, [], batch=1, slots=1, pinned=False)
    schedule = F.fast_mode_frames(probe.frame_count, probe.fps, extract_frequency)
    lo, hi = shard.frame_range(rank, world, len(schedule))
    mine = schedule[lo:hi]
    if sub_area == "default":
        sub_area = default_sub_area(probe.h, probe.w)
    results: Dict[int, object] = {}
    feed = None
    if mine:
        feed = FrameFeed(path, mine[0], mine[-1], mine, batch=batch, pinned=pinned, half=half,
                         decoders=default_decoders(world) if decoders is None else decoders)
        lap("feed_setup_s")
        results = run_feed(engine, feed, stats=stats)
    lap("run_feed_s")
    local = []
    for no in sorted(results):
        ls = rawtxt.lines_from_frame_result(no, results[no], characters, sub_area=sub_area, rec_char_type=rec_char_type,
                                            drop_score=drop_score)
        local.append((no, ls))
    lap("lines_s")
    merged = shard.gather_by_frame(local)
    lines = [l for _, ls in merged for l in ls]
    msec = _gather_msec(feed)
    lap("gather_s")
    subs, text = _srt(lines, probe.fps, path, threshold, msec) if write_srt else ([], "")
    lap("srt_s")
    return JobResult(lines, subs, text, len(results), sorted(results), results)


def accurate_mode_job(engine, path: str, characters: Sequence[str], rank: int = 0, world: int = 1, batch: int = 64,
                      sub_area: Optional[Coordinate] = "default", rec_char_type: str = "ch", drop_score: float = 0.75,
                      threshold: float = 0.8, first: int = 1, last: Optional[int] = None, pinned: bool = True,
                      write_srt: bool = True, stats: Optional[dict] = None, decoders: Optional[int] = None) -> JobResult:
    """Accurate mode of the reference (`extract_frame_by_det`, backend/main.py:255-376): the detector looks at EVERY frame and
    the recogniser reads the frames the controller asks for.  Here det + rec run over every frame of this rank's range in
    batches (`vse_run`), the records are gathered by frame number and the reference's decisions are replayed
    (accurate.accurate_mode_tasks) — so the queued tasks, their cached results and the raw.txt lines come out as the
    reference's own loop produces them.  [first, last]: optional stretch of the video (frame numbers stay the video's)."""
    from .charset import ids_to_text
    probe = FrameFeed(path, 1, 0, [], batch=1, slots=1, pinned=False)
    last = probe.frame_count if last is None else min(last, probe.frame_count)
    n = last - first + 1
    lo, hi = shard.frame_range(rank, world, n)
    if sub_area == "default":
        sub_area = default_sub_area(probe.h, probe.w)
    results: Dict[int, object] = {}
    feed = None
    if hi > lo:
        feed = FrameFeed(path, first + lo, first + hi - 1, None, batch=batch, pinned=pinned,
                         decoders=default_decoders(world) if decoders is None else decoders)
        results = run_feed(engine, feed, stats=stats)
    local = [(no, ([q.tolist() for q in r.quads], [(ids_to_text(i, characters), float(s)) for i, s in zip(r.ids, r.rec_scores)]))
             for no, r in sorted(results.items())]
    merged = dict(shard.gather_by_frame(local))
    delivered = max(merged) - first + 1 if merged else 0

    def detect(k):
        return merged[first + k - 1][0] if first + k - 1 in merged else []

    def predict(k):
        quads, rec = merged[first + k - 1]
        return rawtxt.order_like_predict([np.asarray(q, np.float32) for q in quads], rec)

    tasks = accurate.accurate_mode_tasks(n, detect, predict, sub_area, threshold, frames_read=delivered)
    lines: List[str] = []
    for k, dt_box, rec in tasks:
        if dt_box is None:                        # the worker runs predict on the task's own frame (subtitle_ocr.py:29-30)
            dt_box, rec = predict(k)
        lines += rawtxt.frame_lines(first + k - 1, dt_box, rec, sub_area=sub_area, rec_char_type=rec_char_type, drop_score=drop_score)
    msec = _gather_msec(feed)
    subs, text = _srt(lines, probe.fps, path, threshold, msec) if write_srt else ([], "")
    res = JobResult(lines, subs, text, len(results), sorted(results), results)
    res.tasks = tasks
    return res
