"""Golden vectors for the reference-side glue around the predictor calls:

* ``OcrRecogniser.predict``  (backend/tools/ocr.py:24-86): quad -> (xmin, xmax, ymin, ymax), line buckets, in-line x sort;
* ``get_coordinates``        (backend/tools/ocr.py:115-134);
* ``extract_subtitles``      (backend/tools/subtitle_ocr.py:20-83): ROI overflow rate, DROP_SCORE, the raw.txt line format.

Runs the REFERENCE'S OWN functions (imported from /root/reference, unmodified) on seeded random predictor outputs and
writes inputs + outputs to rawtxt_golden.json.  The GUI / Paddle / geometry packages the reference imports at module
level are absent from this image, so they are stubbed here: qfluentwidgets (config items that only hold their default),
paddle / paddleocr / fsplit (never called), and shapely's Polygon restated for convex polygons (Sutherland-Hodgman clip +
shoelace area: exact for the axis-aligned rectangles the reference builds).  Only runs where /root/reference exists; the
tests read the committed JSON.
"""
import json
import os
import sys
import types

import numpy as np

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "rawtxt_golden.json")


def _stub_modules():
    q = types.ModuleType("qfluentwidgets")

    class _Item:
        def __init__(self, group, name, default=None, *a, **k):
            self.value = default

    class _Any:
        def __init__(self, *a, **k):
            pass

    class _QConfig:
        pass

    class _qconfig:
        @staticmethod
        def load(path, cfg):
            return None

    for n in ("ConfigItem", "OptionsConfigItem", "RangeConfigItem"):
        setattr(q, n, _Item)
    for n in ("OptionsValidator", "BoolValidator", "EnumSerializer", "RangeValidator", "ConfigValidator"):
        setattr(q, n, _Any)
    q.QConfig, q.qconfig = _QConfig, _qconfig
    sys.modules["qfluentwidgets"] = q

    class Polygon:                      # convex polygons only (the reference builds axis-aligned rectangles)
        def __init__(self, pts):
            self.pts = [(float(x), float(y)) for x, y in pts]

        @property
        def area(self):
            p = self.pts
            return abs(sum(p[i][0] * p[(i + 1) % len(p)][1] - p[(i + 1) % len(p)][0] * p[i][1] for i in range(len(p)))) / 2 if len(p) >= 3 else 0.0

        @property
        def is_empty(self):
            return len(self.pts) == 0

        def intersection(self, other):
            def inside(p, a, b):
                return (b[0] - a[0]) * (p[1] - a[1]) - (b[1] - a[1]) * (p[0] - a[0]) >= 0

            def cross_pt(p, q_, a, b):
                x1, y1, x2, y2, x3, y3, x4, y4 = *p, *q_, *a, *b
                d = (x1 - x2) * (y3 - y4) - (y1 - y2) * (x3 - x4)
                t = ((x1 - x3) * (y3 - y4) - (y1 - y3) * (x3 - x4)) / d
                return (x1 + t * (x2 - x1), y1 + t * (y2 - y1))

            clip = other.pts
            if sum(clip[i][0] * clip[(i + 1) % len(clip)][1] - clip[(i + 1) % len(clip)][0] * clip[i][1] for i in range(len(clip))) < 0:
                clip = clip[::-1]
            out = self.pts
            for i in range(len(clip)):
                a, b = clip[i], clip[(i + 1) % len(clip)]
                inp, out = out, []
                for j in range(len(inp)):
                    p, q_ = inp[j], inp[(j + 1) % len(inp)]
                    if inside(q_, a, b):
                        if not inside(p, a, b):
                            out.append(cross_pt(p, q_, a, b))
                        out.append(q_)
                    elif inside(p, a, b):
                        out.append(cross_pt(p, q_, a, b))
                if not out:
                    break
            return Polygon(out)

    sh = types.ModuleType("shapely")
    shg = types.ModuleType("shapely.geometry")
    shg.Polygon = Polygon
    sh.geometry = shg
    sys.modules["shapely"], sys.modules["shapely.geometry"] = sh, shg

    pd = types.ModuleType("paddle")
    pd.is_compiled_with_cuda = lambda: False
    pd.static = types.SimpleNamespace(cuda_places=lambda: [])
    sys.modules["paddle"] = pd
    po = types.ModuleType("paddleocr")
    po.PaddleOCR = object
    sys.modules["paddleocr"] = po
    fs = types.ModuleType("fsplit")
    fsf = types.ModuleType("fsplit.filesplit")
    fsf.Filesplit = object
    fs.filesplit = fsf
    sys.modules["fsplit"], sys.modules["fsplit.filesplit"] = fs, fsf


def main():
    _stub_modules()
    sys.path.insert(0, REF)
    cwd = os.getcwd()
    os.chdir(REF)                       # config.py reads relative paths
    try:
        from backend.tools.ocr import OcrRecogniser, get_coordinates
        from backend.tools import subtitle_ocr as so
        from backend.bean.subtitle_area import SubtitleArea
    finally:
        os.chdir(cwd)
    import tqdm as _tqdm
    so.tqdm.write = lambda *a, **k: None          # the reference logs every line through tqdm.write
    if "Main" not in so.tr:                       # the stubbed config has no interface language: load the English strings
        so.tr.read(os.path.join(REF, "backend", "interface", "en.ini"), encoding="utf-8")
    rng = np.random.default_rng(20260117)
    words = ["As", "far", "as", "we", "can", "go.", "Yami", "Sukehiro", "字幕", "提取", "テスト", "한국어", "x", "Hello,", "world!"]
    cases = []
    for c in range(60):
        n = int(rng.integers(0, 7))
        quads, rec = [], []
        base_y = int(rng.integers(500, 640))
        for k in range(n):
            line = int(rng.integers(0, 3))
            x0 = int(rng.integers(40, 900))
            w, h = int(rng.integers(40, 600)), int(rng.integers(18, 60))
            y0 = base_y + line * int(rng.integers(28, 70)) + int(rng.integers(-6, 7))
            j = rng.integers(-3, 4, size=8)
            q = [[x0 + j[0], y0 + j[1]], [x0 + w + j[2], y0 + j[3]], [x0 + w + j[4], y0 + h + j[5]], [x0 + j[6], y0 + h + j[7]]]
            quads.append(np.asarray(q, np.float32))
            rec.append((" ".join(rng.choice(words, size=int(rng.integers(1, 5)))), float(np.round(rng.uniform(0.3, 1.0), 4))))
        rec_type = ["en", "ch"][c % 2]
        area = None if c % 5 == 4 else dict(ymin=int(rng.integers(450, 560)), ymax=int(rng.integers(660, 720)),
                                            xmin=int(rng.integers(0, 120)), xmax=int(rng.integers(1000, 1280)))
        opts = dict(REC_CHAR_TYPE=rec_type, DROP_SCORE=float(rng.choice([0.0, 0.5, 0.75])),
                    SUB_AREA_DEVIATION_RATE=float(rng.choice([0.0, 0.05, 0.2])), DEBUG_OCR_LOSS=False)
        o = OcrRecogniser.__new__(OcrRecogniser)
        o.recogniser = lambda image, cls=False, _q=quads, _r=rec: (list(_q), list(_r), {})
        dt_box, res = o.predict(None)
        coords = get_coordinates(dt_box)
        raw = []
        sub_area = SubtitleArea(area["ymin"], area["ymax"], area["xmin"], area["xmax"]) if area else None
        so.extract_subtitles({"i": 17 + c}, None, None, raw, sub_area, types.SimpleNamespace(**opts), dt_box, res, "/tmp/none")
        cases.append(dict(frame_no=17 + c, quads=[q.tolist() for q in quads], rec=[[t, s] for t, s in rec], sub_area=area, options=opts,
                          predict_boxes=[[list(p) for p in b] for b in dt_box], predict_res=[[t, s] for t, s in res],
                          coordinates=[list(x) for x in coords], raw_lines=raw))
    with open(OUT, "w", encoding="utf-8") as f:
        json.dump(dict(generator="tests/golden/make_rawtxt_golden.py", reference_functions=[
            "backend/tools/ocr.py:24-86 OcrRecogniser.predict", "backend/tools/ocr.py:115-134 get_coordinates",
            "backend/tools/subtitle_ocr.py:20-83 extract_subtitles"], cases=cases), f, ensure_ascii=False, indent=0)
    print(len(cases), "cases,", sum(len(c["raw_lines"]) for c in cases), "raw.txt lines ->", OUT)


if __name__ == "__main__":
    main()
