"""Per-step error of a packed plan on the engine vs the CPU plan interpreter (oracle), for several precision modes — to see
where a mode loses accuracy.  usage (GPU box): python tools/gpu_step_errors.py V4/ch_det fp32_tc,fp32 [h w]"""
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests.nn_compare import compare_all  # noqa: E402
from video_subtitle_extractor_b200 import engine as E, plan as P, weights  # noqa: E402
from video_subtitle_extractor_b200.synth import SynthStream  # noqa: E402

name = sys.argv[1]
modes = sys.argv[2].split(",")
h, w = (int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else (352, 480)
blob = weights.load_plan_blob(name)
pl = P.deserialize(blob)
which = E.PLAN_DET if weights.is_det(name) else E.PLAN_REC
vid = os.path.join(ROOT, "tests", "golden", "_videos", "test_cn.mp4")
if os.path.exists(vid):
    cap = cv2.VideoCapture(vid)
    cap.set(cv2.CAP_PROP_POS_FRAMES, 134)
    frame = cap.read()[1]
else:
    frame = SynthStream(1080, 1920).frame(0)
img = cv2.resize(frame, (w, h)) if which == E.PLAN_DET else np.ascontiguousarray(frame[int(frame.shape[0] * 0.86):int(frame.shape[0] * 0.86) + 48, 300:940])
PREC = {"fp16": E.PRECISION_FP16, "fp32": E.PRECISION_FP32, "tf32": E.PRECISION_TF32, "fp32_tc": E.PRECISION_FP32_TC}
reps = {}
for m in modes:
    eng = E.Engine(precision=PREC[m])
    eng.load_plan(which, blob, name)
    reps[m] = compare_all(eng, which, pl, [img])
    eng.close()
print("step op            K      " + "  ".join(f"{m:>10s}" for m in modes) + "   (max abs err / max |ref|)")
for i, row in enumerate(reps[modes[0]]):
    k, op, vid_, err, mx = row
    s = pl.steps[k]
    K = s.p.get("kh", 1) * s.p.get("kw", 1) * s.p.get("cin", 0) if op in ("CONV", "DWCONV") else 0
    vals = [reps[m][i][3] / max(reps[m][i][4], 1e-6) for m in modes]
    if max(vals) > 3e-6 or op == "CONV":
        print(f"{k:4d} {op:10s} {K:6d}  " + "  ".join(f"{v:10.2e}" for v in vals))
