"""smoke(): one tiny det+rec invocation of the hot path on cuda:0 through the C-ABI, checked against the CPU oracle."""
from __future__ import annotations

import numpy as np


def run() -> None:
    from oracle.pipeline import OraclePipeline            # the checker (test infrastructure), never the product path
    from . import engine as E
    from . import weights
    from .synth import SynthStream

    if E.device_count() == 0:
        raise RuntimeError("smoke(): no CUDA device (this engine has no CPU fallback)")
    det_blob, rec_blob = weights.load_plan_blob("V4/ch_det_fast"), weights.load_plan_blob("V4/en_rec_fast")
    frames = [SynthStream(540, 960).frame(i) for i in (0, 50)]
    eng = E.Engine(device=0)
    eng.load_plan(E.PLAN_DET, det_blob, "V4/ch_det_fast")
    eng.load_plan(E.PLAN_REC, rec_blob, "V4/en_rec_fast")
    got = eng.run(frames)
    launches = eng.launch_count
    eng.close()
    oracle = OraclePipeline.from_plans(det_blob, rec_blob)
    for f, g in zip(frames, got):
        r = oracle.ocr(f)
        assert len(g.quads) == len(r.boxes), "box count differs from the oracle"
        for q, b in zip(g.quads, r.boxes):
            assert np.abs(q - np.asarray(b, np.float32)).max() <= 2.0, "box differs from the oracle"
        assert g.ids == r.ids, "recognised class ids differ from the oracle"
    assert launches > 100
    print(f"smoke ok: {sum(len(g.quads) for g in got)} text line(s), {launches} kernel launches, ids match the CPU oracle")
