"""Generates tests/golden/*.png + golden.json.  Run HERE (needs /root/reference); the outputs are committed.

Frames come from the reference's own sample videos (reference test/test_en.mp4, test/test_cn.mp4 — the only fixtures the
reference ships for this path); expected results come from the graph-level CPU oracle
(oracle/graph_interp.py executing the shipped inference.pdmodel + oracle/hostlogic.py), i.e. NOT from the plan
compiler and NOT from the CUDA engine.  To stay small the frames are cut to their lower half — which is what the
reference's frame_preprocess does for a bottom subtitle area (backend/tools/subtitle_ocr.py:270-289).
"""
import json
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.pipeline import OraclePipeline  # noqa: E402

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))
CASES = [("test_en.mp4", [300, 600, 900, 1200, 1500, 1800, 2100, 2500, 3000, 3300]), ("test_cn.mp4", [600, 1000])]


def main():
    orc = OraclePipeline(f"{REF}/backend/models/V4/ch_det_fast", f"{REF}/backend/models/V4/en_rec_fast")
    golden = {"models": ["V4/ch_det_fast", "V4/en_rec_fast"], "cases": []}
    for video, frames in CASES:
        cap = cv2.VideoCapture(f"{REF}/test/{video}")
        for fno in frames:
            cap.set(cv2.CAP_PROP_POS_FRAMES, fno)
            ok, frame = cap.read()
            assert ok
            frame = np.ascontiguousarray(frame[frame.shape[0] // 2:])
            name = f"{video.split('.')[0]}_{fno}.png"
            cv2.imwrite(os.path.join(OUT, name), frame, [cv2.IMWRITE_PNG_COMPRESSION, 9])
            frame = cv2.imread(os.path.join(OUT, name))
            res = orc.ocr(frame)
            det = orc.detect(frame)
            pm, _ = orc.det_prob_map(frame)
            golden["cases"].append({
                "image": name, "shape": list(frame.shape),
                "det_only_boxes": det.astype(int).tolist(),
                "boxes": [np.asarray(b).astype(int).tolist() for b in res.boxes],
                "det_scores": [float(s) for s in res.det_scores],
                "ids": res.ids, "rec_scores": [float(s) for s in res.scores], "rec_widths": res.rec_widths,
                "prob_gt_0.3": int((pm > 0.3).sum()), "prob_near_thresh": int(((pm > 0.29) & (pm < 0.31)).sum()),
            })
            print(name, golden["cases"][-1]["boxes"], golden["cases"][-1]["ids"])
    with open(os.path.join(OUT, "golden.json"), "w") as f:
        json.dump(golden, f, indent=1)


if __name__ == "__main__":
    main()
