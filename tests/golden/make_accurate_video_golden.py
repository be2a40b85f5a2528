"""Generates tests/golden/accurate_video_golden_test_cn.json — BASELINE configs[2] (test/test_cn.mp4, accurate mode,
V4/ch_det + V4/ch_rec) on a stretch of the video, entirely from the reference side:

 stage 1  per-frame predictor outputs of EVERY frame of the stretch from the graph-level CPU oracle (the shipped server
          graphs executed op by op, ~12 s per frame on 8 cores — hence a stretch, not the whole video);
 stage 2  the REFERENCE'S OWN accurate-mode loop `SubtitleExtractor.extract_frame_by_det` (+ `_compare_ocr_result`,
          `__get_area_text`, backend/main.py:255-376, 905-952) driven by those outputs through its own
          `OcrRecogniser.predict`, its worker's `extract_subtitles`, `_remove_duplicate_subtitle` and `generate_subtitle_file`
          -> queued tasks, raw.txt lines, .srt text.

Run HERE (needs /root/reference).  tests/test_gpu_jobs.py runs job.accurate_mode_job on the engine and compares."""
import json
import os
import sys
import tempfile
import time
import types

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
VIDEO, DET, REC, FIRST, LAST = "test_cn.mp4", "V4/ch_det", "V4/ch_rec", 41, 136
OUT = os.path.join(HERE, "accurate_video_golden_test_cn.json")


def stage1():
    from oracle.pipeline import OraclePipeline
    orc = OraclePipeline(f"{REF}/backend/models/{DET}", f"{REF}/backend/models/{REC}")
    cap = cv2.VideoCapture(f"{REF}/test/{VIDEO}")
    out = dict(video=VIDEO, models=[DET, REC], first=FIRST, last=LAST, fps=cap.get(cv2.CAP_PROP_FPS),
               frame_count=int(cap.get(cv2.CAP_PROP_FRAME_COUNT)), frames=[])
    no, t0 = 0, time.time()
    while no < LAST:
        ok, frame = cap.read()
        assert ok
        no += 1
        if no < FIRST:
            continue
        r = orc.ocr(frame)
        out["frames"].append(dict(no=no, shape=list(frame.shape), sum=int(frame.sum(dtype=np.uint64)),
                                  boxes=[np.asarray(b).astype(int).tolist() for b in r.boxes],
                                  det_scores=[round(float(s), 6) for s in r.det_scores], ids=r.ids,
                                  rec_scores=[round(float(s), 6) for s in r.scores], rec_widths=r.rec_widths))
        print(no, f"{time.time() - t0:.0f}s", [len(i) for i in r.ids], flush=True)
        with open(OUT + ".stage1", "w") as f:
            json.dump(out, f)
    return out


def stage2(g):
    from video_subtitle_extractor_b200 import charset
    from video_subtitle_extractor_b200.job import default_sub_area
    sys.path.insert(0, HERE)
    import make_dedup_golden as d
    import make_rawtxt_golden as rg
    rg._stub_modules()
    lev = types.ModuleType("Levenshtein")
    lev.ratio = d.indel_ratio
    sys.modules["Levenshtein"] = lev
    from unittest.mock import MagicMock
    for name in ["pysrt", "wordsegment", "imageio_ffmpeg", "onnxruntime"]:
        sys.modules[name] = MagicMock()
    sys.path.insert(0, REF)
    sys.path.insert(0, os.path.join(REF, "backend"))
    cwd = os.getcwd()
    os.chdir(REF)
    try:
        import backend.main as m
        from backend.tools.ocr import OcrRecogniser
        from backend.tools import subtitle_ocr as so
        from backend.bean.subtitle_area import SubtitleArea
    finally:
        os.chdir(cwd)
    m.tqdm = lambda *a, **k: types.SimpleNamespace(update=lambda n: None)
    so.tqdm.write = lambda *a, **k: None
    for tr in (so.tr, m.tr):
        if "Main" not in tr:
            tr.read(os.path.join(REF, "backend", "interface", "en.ini"), encoding="utf-8")
    chars = charset.characters("ch", None, 6625)
    h, w = g["frames"][0]["shape"][:2]
    area = default_sub_area(h, w)
    sub_area = SubtitleArea(area[2], area[3], area[0], area[1])
    n = len(g["frames"])
    by_k = {k + 1: fr for k, fr in enumerate(g["frames"])}      # the loop numbers the stretch's frames from 1
    state = dict(pos=0)

    class Cap:
        def isOpened(self):
            return True

        def read(self):
            if state["pos"] >= n:
                return False, None
            state["pos"] += 1
            return True, state["pos"]

        def release(self):
            pass

    def detect(k):
        qs = by_k[k]["boxes"]
        return (np.asarray(qs, np.float32) if qs else np.zeros((0,), np.float32)), 0.0

    def predict(k):
        fr = by_k[k]
        quads = [np.asarray(q, np.float32) for q in fr["boxes"]]
        rec = [(charset.ids_to_text(i, chars), s) for i, s in zip(fr["ids"], fr["rec_scores"])]
        o = OcrRecogniser.__new__(OcrRecogniser)
        o.recogniser = lambda image, cls=False: (list(quads), list(rec), {})
        return o.predict(None)

    tasks = []
    fake = types.SimpleNamespace(frame_count=n, ocr=types.SimpleNamespace(predict=predict), video_cap=Cap(),
                                 sub_detector=types.SimpleNamespace(detect_subtitle=detect), sub_area=sub_area,
                                 subtitle_ocr_task_queue=types.SimpleNamespace(put=lambda t: tasks.append(t)),
                                 update_progress=lambda **k: None)
    fake._SubtitleExtractor__get_area_text = lambda r, _f=fake: m.SubtitleExtractor._SubtitleExtractor__get_area_text(_f, r)
    fake._compare_ocr_result = lambda *a, _f=fake: m.SubtitleExtractor._compare_ocr_result(_f, *a)
    m.SubtitleExtractor.extract_frame_by_det(fake)
    # the worker: extract_subtitles per queued task (cached results forwarded, else predict on the task's frame)
    raw = []
    opts = types.SimpleNamespace(REC_CHAR_TYPE="ch", DROP_SCORE=0.75, SUB_AREA_DEVIATION_RATE=0.0, DEBUG_OCR_LOSS=False)
    recogniser = types.SimpleNamespace(predict=lambda img: predict(img))
    for t in tasks:
        k = t[1]
        so.extract_subtitles({"i": g["first"] + k - 1}, recogniser, k, raw, sub_area, opts, t[2], t[3], "/tmp/none")
    with tempfile.NamedTemporaryFile("w", suffix=".txt", delete=False, encoding="utf-8") as f:
        f.writelines(raw)
        raw_path = f.name
    srt_path = raw_path + ".srt"
    fx = types.SimpleNamespace(raw_subtitle_path=raw_path, use_vsf=False, subtitle_output_path=srt_path,
                               video_path=f"{REF}/test/{g['video']}", fps=g["fps"], append_output=lambda *a, **k: None)
    fx._concat_content_with_same_frameno = lambda: m.SubtitleExtractor._concat_content_with_same_frameno(fx)
    fx._remove_duplicate_subtitle = lambda: m.SubtitleExtractor._remove_duplicate_subtitle(fx)
    fx._frame_to_timecode = lambda no: m.SubtitleExtractor._frame_to_timecode(fx, no)
    m.SubtitleExtractor.generate_subtitle_file(fx)
    with open(srt_path, encoding="utf-8") as f:
        srt = f.read()
    os.unlink(raw_path)
    os.unlink(srt_path)
    g["area"] = list(area)
    g["tasks"] = [dict(frame_no=t[1], cached=t[2] is not None) for t in tasks]
    g["raw_lines"] = raw
    g["srt"] = srt
    with open(OUT, "w", encoding="utf-8") as f:
        json.dump(g, f, ensure_ascii=False, separators=(",", ":"))
    print(len(tasks), "tasks,", len(raw), "raw lines,", srt.count(" --> "), "subtitles ->", OUT)


if __name__ == "__main__":
    if os.path.exists(OUT + ".stage1") and "--redo" not in sys.argv:
        with open(OUT + ".stage1") as f:
            g = json.load(f)
        if len(g["frames"]) < LAST - FIRST + 1:
            g = stage1()
    else:
        g = stage1()
    stage2(g)
