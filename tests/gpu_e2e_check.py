"""Manual GPU bring-up script (not a pytest file): vse_run on synthetic frames vs the CPU oracle."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from video_subtitle_extractor_b200 import engine as E, weights
from video_subtitle_extractor_b200.synth import SynthStream
from oracle.pipeline import OraclePipeline
from oracle import hostlogic as hl


def main():
    H, W = (1080, 1920) if "--1080" in sys.argv else (540, 960)
    s = SynthStream(H, W)
    idx = [0, 10, 50, 60, 120, 185]
    frames = [s.frame(i) for i in idx]
    det_blob, rec_blob = weights.load_plan_blob("V4/ch_det_fast"), weights.load_plan_blob("V4/en_rec_fast")
    orc = OraclePipeline.from_plans(det_blob, rec_blob)
    t = time.time()
    ref = [orc.ocr(f) for f in frames]
    print(f"oracle: {time.time()-t:.2f}s for {len(frames)} frames")
    for prec in (E.PRECISION_FP32, E.PRECISION_FP16):
        eng = E.Engine(precision=prec)
        eng.load_plan(0, det_blob, "det")
        eng.load_plan(1, rec_blob, "rec")
        t = time.time()
        got = eng.run(frames)
        dt = time.time() - t
        got = eng.run(frames)
        print(f"== prec={'fp32' if prec else 'fp16'} first call {dt*1e3:.1f} ms, timings {eng.last_timings.round(2).tolist()} launches {eng.launch_count}")
        for k, (g, r) in enumerate(zip(got, ref)):
            gb = [q.astype(int).tolist() for q in g.quads]
            rb = [np.asarray(b).astype(int).tolist() for b in r.boxes]
            ok_box = gb == rb
            ok_ids = g.ids == r.ids
            print(f"frame {idx[k]}: boxes {'OK' if ok_box else 'DIFF'} ids {'OK' if ok_ids else 'DIFF'} "
                  f"n={len(gb)}/{len(rb)} widths {g.rec_widths.tolist()} vs {r.rec_widths} truth={s.truth(idx[k])}")
            if not ok_box:
                print("   got", gb, "\n   ref", rb)
            for i in range(min(len(g.ids), len(r.ids))):
                tg, tr = hl.ids_to_text(g.ids[i], hl.EN_CHARACTERS), hl.ids_to_text(r.ids[i], hl.EN_CHARACTERS)
                flag = "" if g.ids[i] == r.ids[i] else "   <<<<"
                print(f"   '{tg}' ({g.rec_scores[i]:.4f}, det {g.det_scores[i]:.4f}) | ref '{tr}' ({r.scores[i]:.4f}, det {r.det_scores[i]:.4f}){flag}")
        det = eng.run(frames, det_only=True)
        dref = [orc.detect(f) for f in frames]
        print("det_only order/box parity:", [np.array_equal(d.quads, r) for d, r in zip(det, dref)])
        eng.close()


if __name__ == "__main__":
    main()
