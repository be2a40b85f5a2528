"""Manual GPU bring-up script: tcgen05 conv path vs the CUDA-core path vs the CPU plan interpreter, step by step."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, cv2
from video_subtitle_extractor_b200 import engine as E, weights, plan as P
from video_subtitle_extractor_b200.synth import SynthStream
from tests.nn_compare import compare_all

def main():
    s = SynthStream(540, 960)
    frames = [s.frame(0), s.frame(60)]
    for which, name in ((0, "V4/ch_det_fast"), (1, "V4/en_rec_fast")):
        blob = weights.load_plan_blob(name)
        pl = P.deserialize(blob)
        if which == 0:
            imgs, vw = [cv2.resize(f, (480, 288)) for f in frames], None
        else:
            base = frames[0][430:478, 200:760]
            imgs = [np.ascontiguousarray(base[:, :w]) for w in (320, 403, 560)]
            vw = [300, 403, 501]
        reps = {}
        for label, flags in (("simt", E.FLAG_NO_TENSOR_CORES), ("tc", 0)):
            eng = E.Engine(precision=E.PRECISION_FP16, flags=flags)
            eng.load_plan(which, blob, name)
            reps[label] = compare_all(eng, which, pl, imgs, vw)
            print(f"== {name} {label}: launches {eng.launch_count} tc_launches {eng.tc_launch_count}")
            eng.close()
        worst = {"simt": 0.0, "tc": 0.0}
        for (k, op, vid, e1, mx), (_, _, _, e2, _) in zip(reps["simt"], reps["tc"]):
            r1, r2 = e1 / max(mx, 1e-6), e2 / max(mx, 1e-6)
            worst["simt"] = max(worst["simt"], r1); worst["tc"] = max(worst["tc"], r2)
            flag = "   <<<<<" if (not np.isfinite(r2)) or r2 > 5e-2 else ""
            if op in ("CONV",) or flag:
                print(f"  step {k:3d} {op:9s} v{vid:<4d} simt rel {r1:.2e}  tc rel {r2:.2e}{flag}")
        print("   worst rel", worst)

if __name__ == "__main__":
    main()
