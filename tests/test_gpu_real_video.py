"""GPU parity on the reference's own sample videos: the WHOLE fast-mode schedule (`extract_frame_by_fps`, reference
backend/main.py:228-251) of test/test_en.mp4 (400 frames, 678 text boxes) and test/test_cn.mp4 through `vse_run`, against
golden vectors from the graph-level CPU oracle (tests/golden/make_video_golden.py — the shipped inference.pdmodel executed
op by op, NOT the plan compiler and NOT the engine).

Bars (BASELINE.json north_star): every box IoU >= 0.99 against its golden box, CER <= 1e-3 on the class-id sequences — in
the mode bench.py times (PARITY_MODE below = what engine.bench_mode() returns).  The per-mode statistics (boxes that are not
integer-identical, IoU < 0.99, CER, score error) are written to gpurun_out/real_video_parity.json for profiles/.

The videos are copies made by __graft_entry__.build() (tests/golden/_videos/, git-ignored, they travel with the snapshot);
the decoded frames are checked against the pixel sums recorded with the golden vectors.
"""
import json
import os

import cv2
import numpy as np
import pytest

from video_subtitle_extractor_b200 import engine as E
from video_subtitle_extractor_b200 import weights

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")
VIDEOS = os.path.join(GOLDEN, "_videos")
OUT = os.path.join(os.path.dirname(HERE), "gpurun_out")


def iou_quads(a, b, shape):
    ma, mb = np.zeros(shape, np.uint8), np.zeros(shape, np.uint8)
    cv2.fillPoly(ma, [np.asarray(a, np.int32)], 1)
    cv2.fillPoly(mb, [np.asarray(b, np.int32)], 1)
    return (ma & mb).sum() / max(1, (ma | mb).sum())


def edit_distance(a, b):
    d = list(range(len(b) + 1))
    for i, x in enumerate(a, 1):
        prev, d[0] = d[0], i
        for j, y in enumerate(b, 1):
            prev, d[j] = d[j], min(d[j] + 1, d[j - 1] + 1, prev + (x != y))
    return d[-1]


def scheduled_frames(golden):
    """Sequential decode (like the reference's fast mode), yields (golden record, frame) for the scheduled frames."""
    path = os.path.join(VIDEOS, golden["video"])
    cap = cv2.VideoCapture(path)
    want = {f["no"]: f for f in golden["frames"]}
    last = max(want)
    no = 0
    while no < last:
        ok, frame = cap.read()
        if not ok:
            break
        no += 1
        if no in want:
            assert int(frame.sum(dtype=np.uint64)) == want[no]["sum"], f"decoder output differs from the golden run at frame {no}"
            yield want[no], frame
    cap.release()


def run_schedule(golden, eng, batch=32):
    st = dict(frames=0, boxes=0, count_mismatch=0, not_identical=0, iou_lt_099=0, min_iou=1.0, id_errors=0, id_symbols=0,
              lines_wrong=0, max_rec_score_err=0.0, max_det_score_err=0.0, worst=[])
    pend = []

    def flush():
        if not pend:
            return
        got = eng.run([f for _, f in pend])
        for (g, f), r in zip(pend, got):
            st["frames"] += 1
            st["boxes"] += len(g["boxes"])
            if len(r.quads) != len(g["boxes"]):
                st["count_mismatch"] += 1
                st["worst"].append((g["no"], "count", len(r.quads), len(g["boxes"])))
                continue
            for k, (q, b) in enumerate(zip(r.quads, g["boxes"])):
                if q.astype(int).tolist() != b:
                    st["not_identical"] += 1
                    iou = float(iou_quads(q, b, f.shape[:2]))
                    st["min_iou"] = min(st["min_iou"], iou)
                    if iou < 0.99:
                        st["iou_lt_099"] += 1
                        st["worst"].append((g["no"], round(iou, 4), q.astype(int).tolist(), b))
                else:
                    st["max_det_score_err"] = max(st["max_det_score_err"], abs(float(r.det_scores[k]) - g["det_scores"][k]))
                e = edit_distance(r.ids[k], g["ids"][k])
                st["id_errors"] += e
                st["id_symbols"] += len(g["ids"][k])
                st["lines_wrong"] += e > 0
                if e == 0:
                    st["max_rec_score_err"] = max(st["max_rec_score_err"], abs(float(r.rec_scores[k]) - g["rec_scores"][k]))
        pend.clear()

    for g, frame in scheduled_frames(golden):
        pend.append((g, frame))
        if len(pend) == batch:
            flush()
    flush()
    st["cer"] = st["id_errors"] / max(1, st["id_symbols"])
    return st


CASES = [("test_en", "V4/ch_det_fast", "V4/en_rec_fast"), ("test_cn", "V4/ch_det_fast", "V4/ch_rec_fast")]
MODES = [("bench", None), ("fp32", E.PRECISION_FP32), ("fp32_tc", E.PRECISION_FP32_TC), ("fp16", E.PRECISION_FP16),
         # detector in the fp32 tensor-core mode (boxes), recogniser with fp16 activations (class ids)
         ("mixed", dict(precision=E.PRECISION_FP16, flags=E.FLAG_DET_FP32_TC))]
# modes that are measured and recorded but not held to the bar: fp16 storage moves the 0.3 threshold crossing of the detector
# (DESIGN.md §5); "mixed" keeps every box and every en line but flips ~3 % of the 6625-class Chinese lines (CER 4.6e-3 on
# test_cn, profiles/r02_real_video_parity.json) — fp16 logits cannot separate near-tied classes of the big dictionary
REPORT_ONLY = {"fp16", "mixed"}


@pytest.mark.parametrize("mode,prec", MODES)
@pytest.mark.parametrize("video,det,rec", CASES)
def test_whole_fast_mode_schedule_matches_graph_oracle(video, det, rec, mode, prec):
    gpath = os.path.join(GOLDEN, f"video_golden_{video}.json")
    with open(gpath) as f:
        golden = json.load(f)
    assert golden["models"] == [det, rec]
    if not os.path.exists(os.path.join(VIDEOS, golden["video"])):
        pytest.skip("tests/golden/_videos/ is absent (run __graft_entry__.build() where the reference tree exists)")
    if not (weights.have_plan(det) and weights.have_plan(rec)):
        pytest.skip("packed plans not present on this machine")
    kw = E.bench_mode() if prec is None else prec if isinstance(prec, dict) else dict(precision=prec)
    eng = E.Engine(**kw)
    eng.load_plan(E.PLAN_DET, weights.load_plan_blob(det), det)
    eng.load_plan(E.PLAN_REC, weights.load_plan_blob(rec), rec)
    st = run_schedule(golden, eng)
    eng.close()
    os.makedirs(OUT, exist_ok=True)
    rec_path = os.path.join(OUT, "real_video_parity.json")
    allr = {}
    if os.path.exists(rec_path):
        with open(rec_path) as f:
            allr = json.load(f)
    allr[f"{video}:{mode}"] = dict(st, engine=kw, models=[det, rec])
    with open(rec_path, "w") as f:
        json.dump(allr, f, indent=1)
    print(f"\n{video} [{mode}]: {st['frames']} frames, {st['boxes']} boxes, count mismatch {st['count_mismatch']}, not identical "
          f"{st['not_identical']}, IoU<0.99 {st['iou_lt_099']} (min {st['min_iou']:.4f}), CER {st['cer']:.2e} ({st['lines_wrong']} lines), "
          f"score err det {st['max_det_score_err']:.1e} rec {st['max_rec_score_err']:.1e}")
    assert st["frames"] == len(golden["frames"]) >= 150
    if mode in REPORT_ONLY:
        return
    assert st["count_mismatch"] == 0, st["worst"][:5]
    assert st["iou_lt_099"] == 0, st["worst"][:5]          # box-for-box, IoU >= 0.99
    assert st["cer"] <= 1e-3                                 # recognised text
