"""BASELINE configs[0] as plumbing on the CPU: a stretch of the reference's test/test_en.mp4 in fast mode with the default
subtitle area -> frame schedule (frames.py) -> predictor (here the CPU oracle; on a B200 box the engine) -> reading order,
ROI filter and raw.txt lines (rawtxt.py) -> de-dup and .srt text (dedup.py).  Checks that the host modules either side of
the predictor call fit together and produce the subtitle the video shows (SURVEY.md Appendix E anchor), and that the
REFERENCE'S OWN glue — OcrRecogniser.predict, extract_subtitles, _remove_duplicate_subtitle, generate_subtitle_file, run
unmodified in a separate process (tests/golden/ref_glue_runner.py) on the same predictor outputs and the same video —
writes the identical raw.txt lines and the identical .srt text."""
import json
import os
import subprocess
import sys
import tempfile

import cv2
import pytest

from oracle import hostlogic as hl
from oracle.pipeline import OraclePipeline
from video_subtitle_extractor_b200 import dedup, frames, rawtxt, weights

VIDEO = "/root/reference/test/test_en.mp4"


@pytest.mark.skipif(not os.path.exists(VIDEO), reason="reference sample video not present")
def test_fast_mode_stretch_of_test_en_to_srt():
    orc = OraclePipeline.from_plans(weights.load_plan_blob("V4/ch_det_fast"), weights.load_plan_blob("V4/en_rec_fast"))
    cap = cv2.VideoCapture(VIDEO)
    fps, n_frames = cap.get(cv2.CAP_PROP_FPS), int(cap.get(cv2.CAP_PROP_FRAME_COUNT))
    h, w = int(cap.get(cv2.CAP_PROP_FRAME_HEIGHT)), int(cap.get(cv2.CAP_PROP_FRAME_WIDTH))
    # default subtitle area of the reference: fractions 0.78 / 0.99 / 0.05 / 0.95 of the frame (backend/config.py:49)
    area = (int(w * 0.05), int(w * 0.95), int(h * 0.78), int(h * 0.99))
    schedule = [k for k in frames.fast_mode_frames(n_frames, fps, 3) if 270 <= k <= 345]
    assert schedule == list(range(271, 346, 9))
    lines, per_frame = [], []
    for k in schedule:
        cap.set(cv2.CAP_PROP_POS_FRAMES, k - 1)          # what ocr_task_producer does (subtitle_ocr.py:191)
        ok, frame = cap.read()
        assert ok
        r = orc.ocr(frame)
        rec = [(hl.ids_to_text(ids, hl.EN_CHARACTERS), float(s)) for ids, s in zip(r.ids, r.scores)]
        per_frame.append(dict(no=k, quads=[[[float(x), float(y)] for x, y in b] for b in r.boxes], rec=[[t, p] for t, p in rec]))
        dt_box, res = rawtxt.order_like_predict([b for b in r.boxes], rec)
        lines += rawtxt.frame_lines(k, dt_box, res, sub_area=area, rec_char_type="en", drop_score=0.75)
    assert len(lines) >= 5 and all(l.split("\t")[0].isdigit() for l in lines)
    subs = dedup.remove_duplicates(lines, 0.8, use_vsf=False)

    def pos_msec(frame_no):                               # the decoder calls of _frame_to_timecode (main.py:738-742)
        cap.set(cv2.CAP_PROP_POS_FRAMES, frame_no)
        ok, _ = cap.read()
        return cap.get(cv2.CAP_PROP_POS_MSEC) if ok else None

    text, _ = dedup.srt_text(subs, fps, pos_msec)
    cap.release()
    # the recogniser drops some blanks (the reference repairs English spacing afterwards with wordsegment): compare without
    assert [s[2].replace(" ", "") for s in subs] == ["Runawaywithme.\n", "Asfaraswecango.\n"], subs
    blocks = text.strip().split("\n\n")
    assert len(blocks) == 2 and blocks[1].split("\n")[2].replace(" ", "") == "Asfaraswecango."
    start, end = blocks[1].split("\n")[1].split(" --> ")
    assert "00:00:09,000" <= start < end <= "00:00:12,000"    # frames ~280..340 of a 29.97 fps video
    assert "Yami" not in text                              # the title at the top of the frame is outside the subtitle area

    # the same predictor outputs through the REFERENCE'S OWN glue, untouched, in its own process: identical raw.txt and .srt
    with tempfile.TemporaryDirectory() as tmp:
        job = dict(video=VIDEO, fps=fps, area=dict(xmin=area[0], xmax=area[1], ymin=area[2], ymax=area[3]), frames=per_frame,
                   options=dict(REC_CHAR_TYPE="en", DROP_SCORE=0.75, SUB_AREA_DEVIATION_RATE=0.0, DEBUG_OCR_LOSS=False))
        with open(os.path.join(tmp, "in.json"), "w", encoding="utf-8") as f:
            json.dump(job, f, ensure_ascii=False)
        runner = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_glue_runner.py")
        p = subprocess.run([sys.executable, runner, os.path.join(tmp, "in.json"), os.path.join(tmp, "out.json")],
                           capture_output=True, text=True, timeout=600)
        assert p.returncode == 0, p.stderr[-600:]
        with open(os.path.join(tmp, "out.json"), encoding="utf-8") as f:
            ref = json.load(f)
    assert ref["raw_lines"] == lines
    assert ref["srt"] == text
