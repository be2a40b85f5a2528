"""Where does 16-bit activation storage move the DB threshold crossing?  (CPU study, test infrastructure: runs the oracle's
plan interpreter on frames of the reference's sample videos — needs /root/reference.)

For every sampled frame the detector plan runs once in fp32 and once per VARIANT with the engine's storage rounding emulated
(step outputs rounded to fp16 / split fp16 hi+lo, CONV weights rounded to fp16); reported per variant: bitmap pixels that
flip at 0.3, max |dp| and the number of boxes that are not integer-identical / have IoU < 0.99 against the fp32 result.

usage: python tools/precision_study.py [video] [n_frames] [variant,variant,...]
"""
import os
import sys
import time

import cv2
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import hostlogic as hl  # noqa: E402
from oracle.plan_interp import PlanInterpreter  # noqa: E402
from video_subtitle_extractor_b200 import plan as P  # noqa: E402
from video_subtitle_extractor_b200 import weights  # noqa: E402
from video_subtitle_extractor_b200.frames import fast_mode_frames  # noqa: E402

REF = "/root/reference"


def r16(t):
    return t.half().float()


def r16x2(t):   # hi + lo fp16 planes (22 significant bits)
    hi = t.half().float()
    return hi + (t - hi).half().float()


def rbf16(t):
    return t.bfloat16().float()


def trunc_tf32(t):   # what kind::tf32 keeps of an fp32 operand: sign, exponent, 10 mantissa bits (low 13 bits dropped)
    return (t.view(torch.int32) & ~0x1FFF).view(torch.float32)


class Tf32x3:
    """Context manager: dense convolutions of the plan interpreter computed as three tf32 products with fp32 accumulation
    (hi*Wh + lo*Wh + hi*Wl, hi = tf32(x), lo = tf32(x - hi)) — the engine's 3xTF32 tensor-core mode; depthwise stays fp32."""

    def __init__(self, terms=3):
        self.terms = terms

    def __enter__(self):
        import torch.nn.functional as F
        import oracle.plan_interp as pi
        self.F, self.orig = pi.F, pi.F.conv2d
        orig, terms = self.orig, self.terms

        def conv2d(x, w, b=None, stride=1, padding=0, dilation=1, groups=1):
            if groups != 1:
                return orig(x, w, b, stride, padding, dilation, groups)
            xh, wh = trunc_tf32(x), trunc_tf32(w)
            y = orig(xh, wh, b, stride, padding)
            if terms >= 2:
                y = y + orig(trunc_tf32(x - xh), wh, None, stride, padding)
            if terms >= 3:
                y = y + orig(xh, trunc_tf32(w - wh), None, stride, padding)
            return y

        pi.F.conv2d = conv2d
        return self

    def __exit__(self, *a):
        self.F.conv2d = self.orig


def make_interp(pl, round_weights, w_fn=r16):
    it = PlanInterpreter(pl)
    if round_weights:
        for s in pl.steps:
            if s.op == P.OP_CONV:
                c = it._wcache[id(s)]
                c["weight"] = w_fn(c["weight"])
    return it


def iou(a, b, shape):
    ma, mb = np.zeros(shape, np.uint8), np.zeros(shape, np.uint8)
    cv2.fillPoly(ma, [np.asarray(a, np.int32)], 1)
    cv2.fillPoly(mb, [np.asarray(b, np.int32)], 1)
    return (ma & mb).sum() / max(1, (ma | mb).sum())


def variants(pl):
    n = len(pl.steps)
    img = [k for k, s in enumerate(pl.steps) if pl.values[s.out].kind == P.KIND_IMG and k < n - 2]   # DB head deconvs are fused (fp32 inside)
    neck0 = next(k for k, s in enumerate(pl.steps) if s.op == P.OP_CONV and s.ins[0] != pl.steps[k - 1].out and k > 10)
    back = [k for k in img if k < neck0]
    neck = [k for k in img if k >= neck0]
    v = {
        "fp16_all": (dict.fromkeys(img, r16), True),
        "fp16_act_only": (dict.fromkeys(img, r16), False),
        "fp16_w_only": ({}, True),
        "fp16_backbone_only": (dict.fromkeys(back, r16), True),
        "fp16_neck_only": (dict.fromkeys(neck, r16), True),
        "bf16_all": (dict.fromkeys(img, rbf16), True),
        "hilo_all": (dict.fromkeys(img, r16x2), False),
        "hilo_neck_fp16_backbone": ({**dict.fromkeys(back, r16), **dict.fromkeys(neck, r16x2)}, True),
        "tf32x1": ({}, False), "tf32x2": ({}, False), "tf32x3": ({}, False),
    }
    return v, img, neck0


def main():
    video = sys.argv[1] if len(sys.argv) > 1 else "test_en.mp4"
    n_frames = int(sys.argv[2]) if len(sys.argv) > 2 else 60
    pl = P.deserialize(weights.load_plan_blob("V4/ch_det_fast"))
    var, img_steps, neck0 = variants(pl)
    names = sys.argv[3].split(",") if len(sys.argv) > 3 else list(var)
    if "per_step" in names:
        names.remove("per_step")
        for k in img_steps:
            var[f"only_step_{k:02d}_{P.OP_NAMES[pl.steps[k].op]}"] = ({k: r16}, False)
            names.append(f"only_step_{k:02d}_{P.OP_NAMES[pl.steps[k].op]}")
    print("neck starts at step", neck0)
    interp = {False: make_interp(pl, False), True: make_interp(pl, True)}
    cap = cv2.VideoCapture(f"{REF}/test/{video}")
    total = int(cap.get(cv2.CAP_PROP_FRAME_COUNT))
    sched = fast_mode_frames(total, cap.get(cv2.CAP_PROP_FPS))
    want = set(sched[:: max(1, len(sched) // n_frames)][:n_frames])
    stats = {nm: dict(flips=0, maxdp=0.0, rms_logit=0.0, nz=0, box_bad=0, box_inexact=0, boxes=0, frames_bad=0) for nm in names}
    no, done, t0 = 0, 0, time.time()
    while done < len(want):
        ok, frame = cap.read()
        if not ok:
            break
        no += 1
        if no not in want:
            continue
        done += 1
        x, shape, _ = hl.det_preprocess(frame)
        xt = torch.from_numpy(np.ascontiguousarray(x))
        p0 = interp[False].run(xt)[0][0, 0].numpy()
        b0, _ = hl.db_postprocess(p0, shape)
        near = (p0 > 0.05) & (p0 < 0.95)
        l0 = np.log(p0[near] / (1 - p0[near]))
        for nm in names:
            hooks, rw = var[nm]
            if nm.startswith("tf32x"):
                with Tf32x3(int(nm[-1])):
                    pv = interp[False].run(xt)[0][0, 0].numpy()
            else:
                pv = interp[rw].run(xt, store_hook=lambda k, s, t: hooks[k](t) if k in hooks else t)[0][0, 0].numpy()
            st = stats[nm]
            st["flips"] += int(((p0 > 0.3) != (pv > 0.3)).sum())
            st["maxdp"] = max(st["maxdp"], float(np.abs(pv - p0).max()))
            pvn = np.clip(pv[near], 1e-6, 1 - 1e-6)
            st["rms_logit"] += float(((np.log(pvn / (1 - pvn)) - l0) ** 2).sum())
            st["nz"] += int(near.sum())
            bv, _ = hl.db_postprocess(pv, shape)
            st["boxes"] += len(b0)
            bad = inexact = 0
            if len(bv) != len(b0):
                bad = abs(len(bv) - len(b0)) + 0
            for a in b0:
                best = max([iou(a, b, frame.shape[:2]) for b in bv], default=0.0)
                if best < 0.99:
                    bad += 1
                if not any(np.array_equal(np.asarray(a).astype(int), np.asarray(b).astype(int)) for b in bv):
                    inexact += 1
            st["box_bad"] += bad
            st["box_inexact"] += inexact
            st["frames_bad"] += 1 if bad else 0
        if done % 10 == 0:
            print(f"{done} frames, {time.time() - t0:.0f}s", flush=True)
    print(f"{video}: {done} frames")
    print(f"{'variant':40s} flips  max|dp|  rms dlogit  boxes  not-identical  IoU<0.99  frames with IoU<0.99")
    for nm in names:
        st = stats[nm]
        print(f"{nm:40s} {st['flips']:5d}  {st['maxdp']:.4f}  {np.sqrt(st['rms_logit'] / max(1, st['nz'])):.5f}    {st['boxes']:5d}  "
              f"{st['box_inexact']:5d}         {st['box_bad']:5d}     {st['frames_bad']:5d}")


if __name__ == "__main__":
    main()
