"""Manual GPU script: V4/ch_det (server detector) in production mode vs the CPU interpreter, per A/B flag."""
import os, sys
import cv2
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.nn_compare import to_bgrx, interp_values
from video_subtitle_extractor_b200 import engine as E, plan as P, weights
from video_subtitle_extractor_b200.synth import SynthStream

DET = "V4/ch_det"
blob = weights.load_plan_blob(DET)
pl = P.deserialize(blob)
out_vid = pl.steps[-1].out
frames = [SynthStream(1080, 1920).frame(i) for i in (0, 60)]
H, W = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (544, 960)
imgs = [cv2.resize(f, (W, H)) for f in frames]
ref = interp_values(pl, imgs, [W, W])[out_vid]
print("interpreter: bitmap px", int((ref > 0.3).sum()))
variants = {"default": 0, "no-tc": E.FLAG_NO_TENSOR_CORES, "no-fast": E.FLAG_NO_FAST_KERNELS, "no-halo": E.FLAG_NO_HALO,
            "no-halo-rowbox": E.FLAG_NO_HALO | E.FLAG_NO_ROWBOX, "no-pack": E.FLAG_NO_PIXEL_PACK,
            "no-gather": E.FLAG_NO_CONCAT_GATHER}
for name, fl in variants.items():
    for keep in (False, True):
        eng = E.Engine(precision=E.PRECISION_FP16, flags=fl)
        eng.load_plan(E.PLAN_DET, blob, DET)
        eng.debug_run_plan(E.PLAN_DET, [to_bgrx(im) for im in imgs], None, keep_all=keep)
        got = eng.debug_get_value(E.PLAN_DET, out_vid)
        d = np.abs(got - ref)
        print(f"{name:16s} keep_all={keep!s:5s} max|d| {d.max():.4f} mean|d| {d.mean():.6f} flips {int(((got > 0.3) != (ref > 0.3)).sum())} nan {int(np.isnan(got).sum())}")
        if keep and name == "default":
            # first step whose error jumps
            vals = interp_values(pl, imgs, [W, W])
            for k, s in enumerate(pl.steps):
                g = eng.debug_get_value(E.PLAN_DET, s.out)
                if g is None or s.out not in vals or g.shape != vals[s.out].shape:
                    continue
                e = np.abs(g - vals[s.out]).max() / max(np.abs(vals[s.out]).max(), 1e-3)
                if e > 0.03:
                    print("   step", k, P.OP_NAMES[s.op], "rel err", round(float(e), 4))
        eng.close()
