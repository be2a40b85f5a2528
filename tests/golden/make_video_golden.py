"""Generates tests/golden/video_golden_<video>.json: the WHOLE fast-mode schedule of the reference's sample videos through
the graph-level CPU oracle.  Run HERE (needs /root/reference); the JSON is committed, the videos themselves are copied by
__graft_entry__.build() into tests/golden/_videos/ (git-ignored: reference data, not history; they travel to the GPU box
with the snapshot like the packed plans do).

* frames: decoded sequentially with cv2 and numbered from 1 exactly like the reference's fast mode
  (`extract_frame_by_fps`, reference backend/main.py:228-251; restated in frames.py::fast_mode_frames);
* expected results: oracle/graph_interp.py executing the reference's shipped inference.pdmodel + oracle/hostlogic.py,
  one FULL frame per call as `OcrRecogniser.predict` does (reference backend/tools/ocr.py:27) — NOT the plan compiler and
  NOT the CUDA engine.

usage: python tests/golden/make_video_golden.py [video det rec [max_tasks]] ...   (no arguments: the default set)
"""
import json
import os
import sys
import time

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.pipeline import OraclePipeline  # noqa: E402
from video_subtitle_extractor_b200.frames import fast_mode_frames  # noqa: E402

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))
# (video, detector, recogniser, max tasks): models as PaddleModelConfig picks them for mode=fast (reference
# backend/tools/paddle_model_config.py:53-55) with language en / ch
DEFAULT = [("test_en.mp4", "V4/ch_det_fast", "V4/en_rec_fast", 0), ("test_cn.mp4", "V4/ch_det_fast", "V4/ch_rec_fast", 0)]


def run(video, det, rec, max_tasks=0, tag=None, stride=1):
    orc = OraclePipeline(f"{REF}/backend/models/{det}", f"{REF}/backend/models/{rec}")
    cap = cv2.VideoCapture(f"{REF}/test/{video}")
    fps = cap.get(cv2.CAP_PROP_FPS)
    total = int(cap.get(cv2.CAP_PROP_FRAME_COUNT))
    wanted = fast_mode_frames(total, fps)[::stride]
    if max_tasks:
        wanted = wanted[:max_tasks]
    want = set(wanted)
    out = {"video": video, "models": [det, rec], "fps": fps, "frame_count": total, "schedule": "fast_mode_frames(frame_count, fps)"
           + (f"[::{stride}]" if stride > 1 else "") + (f"[:{max_tasks}]" if max_tasks else ""), "frames": []}
    no, t0 = 0, time.time()
    while True:
        ok, frame = cap.read()
        if not ok:
            break
        no += 1
        if no not in want:
            continue
        res = orc.ocr(frame)
        out["frames"].append({
            "no": no, "shape": list(frame.shape), "sum": int(frame.sum(dtype=np.uint64)),
            "boxes": [np.asarray(b).astype(int).tolist() for b in res.boxes],
            "det_scores": [round(float(s), 6) for s in res.det_scores],
            "ids": res.ids, "rec_scores": [round(float(s), 6) for s in res.scores], "rec_widths": res.rec_widths})
        if len(out["frames"]) % 20 == 0:
            print(video, len(out["frames"]), "/", len(wanted), f"{time.time() - t0:.0f}s", flush=True)
        if no >= wanted[-1]:
            break
    name = tag or video.split(".")[0]
    with open(os.path.join(OUT, f"video_golden_{name}.json"), "w") as f:
        json.dump(out, f, separators=(",", ":"))
    print(video, "frames", len(out["frames"]), "boxes", sum(len(x["boxes"]) for x in out["frames"]))


if __name__ == "__main__":
    if len(sys.argv) > 1:
        a = sys.argv[1:]
        run(a[0], a[1], a[2], int(a[3]) if len(a) > 3 else 0, a[4] if len(a) > 4 else None, int(a[5]) if len(a) > 5 else 1)
    else:
        for v, d, r, m in DEFAULT:
            run(v, d, r, m)
