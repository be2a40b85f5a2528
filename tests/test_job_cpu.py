"""job.py on the CPU: the frame feed, sharding, raw.txt / de-dup / .srt tail and the accurate-mode replay, driven by a FAKE
engine that answers every frame with the graph-level oracle's recorded result (tests/golden/video_golden_*.json, looked up by
the frame's pixel sum) — so the host plumbing is checked against the reference glue's goldens without a GPU.  The GPU twin
is tests/test_gpu_jobs.py (real engine)."""
import ctypes
import json
import os

import numpy as np
import pytest

from video_subtitle_extractor_b200 import charset, job
from video_subtitle_extractor_b200.engine import FrameResult

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")
VIDEOS = os.path.join(GOLDEN, "_videos")


class FakeEngine:
    def __init__(self, records):
        self.by_sum = {r["sum"]: r for r in records}
        self.prefetched = 0
        self.calls = []

    def prefetch(self, ptrs, hs, ws, st=None, mem_kind=1):
        self.prefetched += 1

    def run_device(self, ptrs, hs, ws, st=None, det_only=False, mem_kind=2):
        out = []
        self.calls.append(len(ptrs))
        for p, h, w, s in zip(ptrs, hs, ws, st):
            buf = (ctypes.c_uint8 * (h * s)).from_address(p)
            frame = np.frombuffer(buf, np.uint8).reshape(h, s)[:, :w * 3]
            r = self.by_sum[int(frame.sum(dtype=np.uint64))]
            n = len(r["boxes"])
            out.append(FrameResult(np.asarray(r["boxes"], np.float32).reshape(n, 4, 2), np.asarray(r["det_scores"], np.float32),
                                   [list(i) for i in r["ids"]], np.asarray(r["rec_scores"], np.float32),
                                   np.asarray(r["rec_widths"], np.int32)))
        return out


def _need(name):
    if not os.path.exists(os.path.join(VIDEOS, name)):
        pytest.skip("tests/golden/_videos/ is absent")
    return os.path.join(VIDEOS, name)


@pytest.mark.parametrize("world,decoders", [(1, 1), (3, 1), (1, 3), (2, 2)])
def test_fast_mode_job_on_recorded_results(world, decoders):
    path = _need("test_en.mp4")
    with open(os.path.join(GOLDEN, "video_golden_test_en.json")) as f:
        vg = json.load(f)
    with open(os.path.join(GOLDEN, "job_golden_test_en.json"), encoding="utf-8") as f:
        g = json.load(f)
    chars = charset.characters("en", None, 97)
    lines, numbers = [], []
    for rank in range(world):       # ranks run one after the other here; gather_by_frame is the identity without a process group
        eng = FakeEngine(vg["frames"])
        # decoders > 1: the rank's schedule is cut into segments with one seeking decoder thread each; the fake engine finds
        # every frame by its pixel sum, so a decoder that landed on the wrong frame after its seek fails the lookup
        res = job.fast_mode_job(eng, path, chars, rank=rank, world=world, batch=32, sub_area=tuple(g["area"]), pinned=False,
                                write_srt=False, decoders=decoders)
        lines += res.lines
        numbers += res.frame_numbers
        assert eng.prefetched == max(len(eng.calls) - 1, 0) and max(eng.calls) <= 32
    assert numbers == [fr["no"] for fr in vg["frames"]]            # contiguous ranges, nothing lost or doubled
    assert lines == g["raw_lines"]                                  # the reference's own extract_subtitles output
    subs, text = job._srt(lines, vg["fps"], path, 0.8)
    assert text == g["srt"]                                         # ... and its own generate_subtitle_file


@pytest.mark.parametrize("decoders", [1, 2])
def test_accurate_mode_job_on_recorded_results(decoders):
    path = _need("test_cn.mp4")
    with open(os.path.join(GOLDEN, "accurate_video_golden_test_cn.json"), encoding="utf-8") as f:
        g = json.load(f)
    eng = FakeEngine(g["frames"])
    res = job.accurate_mode_job(eng, path, charset.characters("ch", None, 6625), batch=64, sub_area=tuple(g["area"]),
                                rec_char_type="ch", first=g["first"], last=g["last"], pinned=False, decoders=decoders)
    assert sorted(eng.calls, reverse=True) == [64, 32]
    assert [(t[0], t[1] is not None) for t in res.tasks] == [(t["frame_no"], t["cached"]) for t in g["tasks"]]
    assert res.lines == g["raw_lines"]
    assert res.srt == g["srt"]


def test_frame_feed_segments_cover_the_request_exactly():
    """More decoders than batches, a ragged last batch, an empty request and a whole-range (wanted=None) feed: every requested
    frame is delivered exactly once, whatever the number of decoder threads."""
    path = _need("test_en.mp4")
    wanted = list(range(100, 400, 7))            # 43 frames: batches of 8 -> 6 batches, the last one short
    for k in (1, 2, 4, 9):
        feed = job.FrameFeed(path, 90, 420, wanted, batch=8, pinned=False, decoders=k)
        got = []
        for b in feed:
            assert 1 <= len(b.numbers) <= 8 and b.numbers == sorted(b.numbers)
            got += b.numbers
            feed.release(b)
        assert sorted(got) == wanted and len(feed._segments) == min(k, 6)
        assert all(n in feed.msec for n in wanted)
    feed = job.FrameFeed(path, 1, 0, [], batch=4, pinned=False, decoders=3)
    assert list(feed) == []
    feed = job.FrameFeed(path, 50, 75, None, batch=8, pinned=False, decoders=3)      # every frame of the range
    got = []
    for b in feed:
        got += b.numbers
        feed.release(b)
    assert sorted(got) == list(range(50, 76))
