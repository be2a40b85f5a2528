"""Golden vectors for the accurate-mode frame selection: the reference's own
``SubtitleExtractor.extract_frame_by_det`` + ``_compare_ocr_result`` + ``__get_area_text`` (backend/main.py:255-376,
905-952) driven by seeded per-frame detector / OCR results (a fake capture, detector and recogniser return the scripted
results of frame k), recording the OCR tasks it queues.  Stubs as in make_dedup_golden.py.  Only runs where
/root/reference exists."""
import json
import os
import sys
import types

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "accurate_golden.json")


def main():
    sys.path.insert(0, HERE)
    import make_dedup_golden as d
    import make_rawtxt_golden as g
    g._stub_modules()
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
        from backend.bean.subtitle_area import SubtitleArea
    finally:
        os.chdir(cwd)
    m.tqdm = lambda *a, **k: types.SimpleNamespace(update=lambda n: None)
    rng = np.random.default_rng(20260120)
    phrases = ["As far as we can go.", "Yami Sukehiro", "Let's get out of here!", "I don't know.", "Where are you going?", "OK", ""]
    area = dict(ymin=560, ymax=715, xmin=60, xmax=1220)
    cases = []
    for c in range(30):
        n_frames = int(rng.integers(5, 90))
        # script: per frame (1-based) a list of (quad, text)
        script = {}
        k = 1
        while k <= n_frames:
            if rng.random() < 0.3:                     # gap without text
                k += int(rng.integers(1, 6))
                continue
            text, run = str(rng.choice(phrases)), int(rng.integers(1, 15))
            inside = rng.random() < 0.85
            for _ in range(run):
                if k > n_frames:
                    break
                t = text
                if rng.random() < 0.2 and len(t) > 3:    # OCR noise
                    p = int(rng.integers(0, len(t)))
                    t = t[:p] + "x" + t[p + 1:]
                x0, y0 = int(rng.integers(100, 300)), (int(rng.integers(590, 600)) if inside else int(rng.integers(20, 60)))
                quad = [[x0, y0], [x0 + 600, y0], [x0 + 600, y0 + 40], [x0, y0 + 40]]
                boxes = [(quad, t)]
                if rng.random() < 0.2:                   # a second box outside the subtitle area (logo)
                    boxes.append(([[30, 20], [200, 20], [200, 50], [30, 50]], "LOGO"))
                script[k] = boxes
                k += 1
        tasks = []
        state = dict(pos=0)

        class Cap:
            def isOpened(self):
                return True

            def read(self):
                if state["pos"] >= n_frames:
                    return False, None
                state["pos"] += 1
                return True, state["pos"]           # the "frame" is its 1-based number

            def release(self):
                pass

        def detect(frame):
            qs = [q for q, _ in script.get(frame, [])]
            return (np.asarray(qs, np.float32) if qs else np.zeros((0,), np.float32)), 0.0

        def predict(frame):
            b = script.get(frame, [])
            dt = [[(q[0][0], q[0][1]), (q[1][0], q[0][1]), (q[1][0], q[2][1]), (q[0][0], q[2][1])] for q, _ in b]
            return dt, [(t, 0.99) for _, t in b]

        use_area = c % 6 != 5
        fake = types.SimpleNamespace(frame_count=n_frames, ocr=types.SimpleNamespace(predict=predict), video_cap=Cap(),
                                     sub_detector=types.SimpleNamespace(detect_subtitle=detect),
                                     sub_area=SubtitleArea(area["ymin"], area["ymax"], area["xmin"], area["xmax"]) if use_area else None,
                                     subtitle_ocr_task_queue=types.SimpleNamespace(put=lambda t: tasks.append(t)),
                                     update_progress=lambda **k: None)
        fake._SubtitleExtractor__get_area_text = lambda r, _f=fake: m.SubtitleExtractor._SubtitleExtractor__get_area_text(_f, r)
        fake._compare_ocr_result = lambda *a, _f=fake: m.SubtitleExtractor._compare_ocr_result(_f, *a)
        m.SubtitleExtractor.extract_frame_by_det(fake)
        cases.append(dict(n_frames=n_frames, sub_area=area if use_area else None,
                          script={str(k): [[q, t] for q, t in v] for k, v in script.items()},
                          tasks=[dict(frame_no=t[1], dt_box=[[list(p) for p in b] for b in t[2]] if t[2] is not None else None,
                                      rec_res=[[x, s] for x, s in t[3]] if t[3] is not None else None) for t in tasks]))
    with open(OUT, "w", encoding="utf-8") as f:
        json.dump(dict(generator="tests/golden/make_accurate_golden.py", threshold=0.8, reference_functions=[
            "backend/main.py:255-376 extract_frame_by_det", "backend/main.py:905-922 __get_area_text",
            "backend/main.py:924-952 _compare_ocr_result"], cases=cases), f, ensure_ascii=False, indent=0)
    print(len(cases), "cases,", sum(len(c["tasks"]) for c in cases), "queued OCR tasks ->", OUT)


if __name__ == "__main__":
    main()
