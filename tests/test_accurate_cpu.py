"""Accurate-mode frame selection (video_subtitle_extractor_b200/accurate.py) against the tasks the reference's own
`extract_frame_by_det` queues for the same scripted per-frame results (tests/golden/accurate_golden.json, written by
tests/golden/make_accurate_golden.py where /root/reference exists)."""
import json
import os

import numpy as np

from video_subtitle_extractor_b200 import accurate

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "accurate_golden.json")


def test_tasks_match_reference_extract_frame_by_det():
    with open(GOLDEN, encoding="utf-8") as f:
        g = json.load(f)
    n_tasks = n_with_result = 0
    for c in g["cases"]:
        script = {int(k): v for k, v in c["script"].items()}

        def detect(k):
            qs = [q for q, _ in script.get(k, [])]
            return np.asarray(qs, np.float32) if qs else np.zeros((0,), np.float32)

        def predict(k):
            b = script.get(k, [])
            dt = [[(q[0][0], q[0][1]), (q[1][0], q[0][1]), (q[1][0], q[2][1]), (q[0][0], q[2][1])] for q, _ in b]
            return dt, [(t, 0.99) for _, t in b]

        a = c["sub_area"]
        area = (a["xmin"], a["xmax"], a["ymin"], a["ymax"]) if a else None
        got = accurate.accurate_mode_tasks(c["n_frames"], detect, predict, area, g["threshold"])
        want = [(t["frame_no"], t["dt_box"], t["rec_res"]) for t in c["tasks"]]
        norm = lambda tasks: [(no, [[list(p) for p in b] for b in dt] if dt is not None else None,
                               [[x, s] for x, s in rec] if rec is not None else None) for no, dt, rec in tasks]
        assert norm(got) == norm(want), (c["n_frames"], [t[0] for t in got], [t[0] for t in want])
        n_tasks += len(got)
        n_with_result += sum(1 for t in got if t[1] is not None)
        if area is None:
            assert got == []                     # the reference never opens a subtitle without an area
    assert n_tasks > 150 and 0 < n_with_result < n_tasks
