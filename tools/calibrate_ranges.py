"""Per-layer activation ranges for the fp32 tensor-core mode (VSE_PRECISION_FP32_TC).

That mode multiplies fp32 activations by a power of two before splitting them into fp16 hi + lo for the tensor core (which
flushes fp16 subnormals): the larger the factor, the more of the lo halves stay normal numbers, but factor * max |x| must
stay below 65504.  This tool runs the CPU plan interpreter (oracle/) over frames of the reference's sample videos and
synthetic frames and records max |x| over the INPUT of every CONV step of a model; the engine turns it into a per-step shift
with 4x headroom (engine.py::load_plan -> vse_set_conv_input_ranges).  Output: video_subtitle_extractor_b200/calibration/
<model>.json (committed; small).  Needs /root/reference for the videos; models without a file run with a conservative default.

usage: python tools/calibrate_ranges.py [model ...]
"""
import json
import os
import sys

import cv2
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import hostlogic as hl  # noqa: E402
from oracle.pipeline import OraclePipeline  # noqa: E402
from oracle.plan_interp import PlanInterpreter  # noqa: E402
from video_subtitle_extractor_b200 import plan as P  # noqa: E402
from video_subtitle_extractor_b200 import weights  # noqa: E402
from video_subtitle_extractor_b200.synth import SynthStream  # noqa: E402

VIDEOS = os.path.join(ROOT, "tests", "golden", "_videos")
OUT = os.path.join(ROOT, "video_subtitle_extractor_b200", "calibration")
# model -> (videos, frames per video)
DET = {"V4/ch_det_fast": (["test_en.mp4", "test_cn.mp4", "test_japan.mp4", "test_korean.flv"], 10), "V4/ch_det": (["test_cn.mp4", "test_en.mp4"], 3)}
REC = {"V4/en_rec_fast": (["test_en.mp4"], 16), "V4/ch_rec_fast": (["test_cn.mp4"], 12), "V4/ch_rec": (["test_cn.mp4"], 4),
       "V3/japan_rec_fast": (["test_japan.mp4"], 10), "V3/korean_rec_fast": (["test_korean.flv"], 10), "V2/ch_rec": (["test_cn.mp4"], 3)}


def frames_of(video, n):
    cap = cv2.VideoCapture(os.path.join(VIDEOS, video))
    total = int(cap.get(cv2.CAP_PROP_FRAME_COUNT))
    for k in range(n):
        cap.set(cv2.CAP_PROP_POS_FRAMES, int((k + 0.5) * total / n))
        ok, f = cap.read()
        if ok:
            yield f
    cap.release()


def synth_frames():
    for h, w in ((1080, 1920), (720, 1280)):
        s = SynthStream(h, w)
        for i in (0, 60, 120, 50):
            yield s.frame(i)


def record(pl, env, mx):
    for k, s in enumerate(pl.steps):
        if s.op in (P.OP_CONV, P.OP_DECONV2) and s.ins[0] in env:
            mx[k] = max(mx.get(k, 0.0), float(env[s.ins[0]].abs().max()))


def calibrate_det(name):
    pl = P.deserialize(weights.load_plan_blob(name))
    it = PlanInterpreter(pl)
    mx = {}
    vids, n = DET[name]
    srcs = [f for v in vids for f in frames_of(v, n)] + (list(synth_frames()) if n > 3 else list(synth_frames())[:2])
    for f in srcs:
        x, _, _ = hl.det_preprocess(f)
        _, env = it.run(torch.from_numpy(np.ascontiguousarray(x)), keep_all=True)
        record(pl, env, mx)
    return pl, mx, len(srcs)


def calibrate_rec(name):
    pl = P.deserialize(weights.load_plan_blob(name))
    it = PlanInterpreter(pl)
    det = OraclePipeline.from_plans(weights.load_plan_blob("V4/ch_det_fast"))
    rec_h = 32 if name.startswith("V2/") else 48
    mx, n_crops = {}, 0
    vids, n = REC[name]
    for f in [f for v in vids for f in frames_of(v, n)] + list(synth_frames())[:4]:
        boxes = det.detect(f)
        if len(boxes) == 0:
            continue
        crops = [hl.get_rotate_crop_image(f, np.array(b, np.float32)) for b in hl.sorted_boxes(boxes)]
        ratios = [c.shape[1] / float(c.shape[0]) for c in crops]
        for idxs, img_w in hl.rec_batches(ratios, 6, rec_h, 320):
            batch = np.stack([hl.resize_norm_img(crops[i], img_w, rec_h) for i in idxs])
            _, env = it.run(torch.from_numpy(batch), keep_all=True)
            record(pl, env, mx)
            n_crops += len(idxs)
    return pl, mx, n_crops


def main():
    names = sys.argv[1:] or list(DET) + list(REC)
    os.makedirs(OUT, exist_ok=True)
    for name in names:
        if not weights.have_plan(name):
            print("no packed plan for", name)
            continue
        pl, mx, n = calibrate_det(name) if name in DET else calibrate_rec(name)
        rec = {"model": name, "samples": n, "n_steps": len(pl.steps),
               "conv_input_absmax": {str(k): round(v, 4) for k, v in sorted(mx.items())}}
        with open(os.path.join(OUT, name.replace("/", "__") + ".json"), "w") as f:
            json.dump(rec, f, indent=0)
        top = sorted(mx.items(), key=lambda kv: -kv[1])[:4]
        print(name, n, "samples; largest conv inputs:", [(k, round(v, 1)) for k, v in top], flush=True)


if __name__ == "__main__":
    main()
