"""Single-frame latency of vse_run (what the unbatched drop-in shim pays per OcrRecogniser.predict call): one 1080p frame from
pageable host memory per call, bench mode, 200 calls after warm-up."""
import sys, time, json
import numpy as np
sys.path.insert(0, ".")
from video_subtitle_extractor_b200 import engine as E, weights
from video_subtitle_extractor_b200.synth import SynthStream

eng = E.Engine(**E.bench_mode())
eng.load_plan(E.PLAN_DET, weights.load_plan_blob("V4/ch_det_fast"), "V4/ch_det_fast")
eng.load_plan(E.PLAN_REC, weights.load_plan_blob("V4/en_rec_fast"), "V4/en_rec_fast")
s = SynthStream(1080, 1920)
frames = [s.frame(i * 7) for i in range(16)]
for f in frames[:8]:
    eng.run([f])
out = {}
for n in (1, 4):
    t0 = time.perf_counter()
    k = 0
    for r in range(200 // n):
        eng.run([frames[(r * n + j) % 16] for j in range(n)])
        k += n
    dt = time.perf_counter() - t0
    out[f"frames_per_call_{n}"] = {"ms_per_call": round(dt / (200 // n) * 1e3, 3), "frames_per_s": round(k / dt, 1)}
print(json.dumps(out))
