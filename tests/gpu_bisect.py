"""Manual GPU script: which specialised path moves the detector's probability map, on the golden real-video frames."""
import json, os, sys
import cv2
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.nn_compare import to_bgrx
from video_subtitle_extractor_b200 import engine as E, plan as P, weights

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
DET = "V4/ch_det_fast"
blob = weights.load_plan_blob(DET)
pl = P.deserialize(blob)
out_vid = pl.steps[-1].out
golden = json.load(open(os.path.join(GOLDEN, "golden.json")))
frames = [cv2.imread(os.path.join(GOLDEN, c["image"])) for c in golden["cases"]]


def det_in(f):
    h, w = f.shape[:2]
    ratio = min(1.0, 960.0 / max(h, w))
    rh = max(int(round(int(h * ratio) / 32) * 32), 32)
    rw = max(int(round(int(w * ratio) / 32) * 32), 32)
    return cv2.resize(f, (rw, rh))


def prob(flags, prec, img):
    eng = E.Engine(precision=prec, flags=flags)
    eng.load_plan(E.PLAN_DET, blob, DET)
    eng.debug_run_plan(E.PLAN_DET, [to_bgrx(img)], None, keep_all=False)
    out = eng.debug_get_value(E.PLAN_DET, out_vid).reshape(img.shape[0], img.shape[1])
    eng.close()
    return out


variants = {"all-fast": 0, "no-fast": E.FLAG_NO_FAST_KERNELS, "simt": E.FLAG_NO_FAST_KERNELS | E.FLAG_NO_TENSOR_CORES,
            "no-head": E.FLAG_NO_FUSED_HEAD, "no-se": E.FLAG_NO_SE_FUSION, "no-rowbox": E.FLAG_NO_ROWBOX,
            "no-dw": E.FLAG_NO_FAST_DW, "no-stem": E.FLAG_NO_FAST_STEM,
            "fast-no-tc": E.FLAG_NO_TENSOR_CORES}
for c, f in zip(golden["cases"], frames):
    img = det_in(f)
    ref = prob(0, E.PRECISION_FP32, img)
    print(c["image"], img.shape, "fp32 bitmap px", int((ref > 0.3).sum()))
    for name, fl in variants.items():
        p = prob(fl, E.PRECISION_FP16, img)
        flips = (p > 0.3) != (ref > 0.3)
        ys, xs = np.nonzero(flips)
        print(f"   {name:10s} max|d| {np.abs(p - ref).max():.4f} mean|d| {np.abs(p - ref).mean():.6f} flips {int(flips.sum())}"
              f" bbox {(int(xs.min()), int(ys.min()), int(xs.max()), int(ys.max())) if len(xs) else None}")
