"""Manual GPU bring-up script (not a pytest file): per-step parity report for the two benchmark plans."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, cv2
from video_subtitle_extractor_b200 import engine as E, weights, plan as P
from video_subtitle_extractor_b200.synth import SynthStream
from tests.nn_compare import compare_all

def main():
    s = SynthStream(540, 960)
    frames = [s.frame(0), s.frame(60)]
    for prec in (E.PRECISION_FP32, E.PRECISION_FP16):
        eng = E.Engine(precision=prec)
        for which, name in ((0, "V4/ch_det_fast"), (1, "V4/en_rec_fast")):
            blob = weights.load_plan_blob(name)
            eng.load_plan(which, blob, name)
            pl = P.deserialize(blob)
            if which == 0:
                imgs = [cv2.resize(f, (480, 288)) for f in frames]
                vw = None
            else:
                base = frames[0][430:478, 200:760]
                imgs = [np.ascontiguousarray(base[:, :w]) for w in (320, 403, 560)]
                vw = [300, 403, 501]
            t = time.time()
            rep = compare_all(eng, which, pl, imgs, vw)
            print(f"== {name} prec={'fp32' if prec else 'fp16'} ({time.time()-t:.1f}s) launches={eng.launch_count}")
            worst = 0
            for k, op, vid, err, mx in rep:
                rel = err / max(mx, 1e-6)
                worst = max(worst, rel if np.isfinite(rel) else 1e9)
                flag = "" if rel < (1e-4 if prec else 2e-2) else "   <<<<<"
                print(f"  step {k:3d} {op:9s} v{vid:<4d} maxerr {err:.3e} refmax {mx:.3e} rel {rel:.2e}{flag}")
            print(f"   worst rel {worst:.3e}")
        # resize parity
        rng = np.random.default_rng(0)
        for (h, w, dh, dw) in [(1080, 1920, 544, 960), (720, 1280, 544, 960), (61, 733, 48, 577), (30, 200, 48, 320)]:
            img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
            got = eng.debug_resize(img, dh, dw)[:, :, :3]
            want = cv2.resize(img, (dw, dh))
            print("resize", (h, w, dh, dw), "mismatch", int((got != want).sum()))
        eng.close()

if __name__ == "__main__":
    main()
