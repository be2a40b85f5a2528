"""Runs the REFERENCE'S OWN glue (OcrRecogniser.predict -> extract_subtitles -> _remove_duplicate_subtitle ->
generate_subtitle_file) on predictor outputs read from a JSON file, in its own process (the stubs for the reference's
missing imports must not leak into the test process).  Used by tests/test_config0_plumbing_cpu.py where /root/reference
exists.  argv: <in.json> <out.json>;  in: {video, fps, area{xmin,xmax,ymin,ymax}, options{...}, frames:[{no, quads, rec}]}"""
import json
import os
import sys
import tempfile
import types

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    src, dst = sys.argv[1], sys.argv[2]
    with open(src, encoding="utf-8") as f:
        job = json.load(f)
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
        from backend.tools.ocr import OcrRecogniser
        from backend.tools import subtitle_ocr as so
        from backend.bean.subtitle_area import SubtitleArea
    finally:
        os.chdir(cwd)
    so.tqdm.write = lambda *a, **k: None
    for tr in (so.tr, m.tr):
        if "Main" not in tr:
            tr.read(os.path.join(REF, "backend", "interface", "en.ini"), encoding="utf-8")
    a = job["area"]
    area = SubtitleArea(a["ymin"], a["ymax"], a["xmin"], a["xmax"])
    raw = []
    for fr in job["frames"]:
        quads = [np.asarray(q, np.float32) for q in fr["quads"]]
        rec = [(t, s) for t, s in fr["rec"]]
        o = OcrRecogniser.__new__(OcrRecogniser)
        o.recogniser = lambda image, cls=False, _q=quads, _r=rec: (list(_q), list(_r), {})
        dt_box, res = o.predict(None)
        so.extract_subtitles({"i": fr["no"]}, None, None, raw, area, types.SimpleNamespace(**job["options"]), dt_box, res, "/tmp/none")
    with tempfile.NamedTemporaryFile("w", suffix=".txt", delete=False, encoding="utf-8") as f:
        f.writelines(raw)
        raw_path = f.name
    srt_path = raw_path + ".srt"
    fake = types.SimpleNamespace(raw_subtitle_path=raw_path, use_vsf=False, subtitle_output_path=srt_path, video_path=job["video"],
                                 fps=job["fps"], append_output=lambda *a, **k: None)
    fake._concat_content_with_same_frameno = lambda: m.SubtitleExtractor._concat_content_with_same_frameno(fake)
    fake._remove_duplicate_subtitle = lambda: m.SubtitleExtractor._remove_duplicate_subtitle(fake)
    fake._frame_to_timecode = lambda no: m.SubtitleExtractor._frame_to_timecode(fake, no)
    m.SubtitleExtractor.generate_subtitle_file(fake)
    with open(srt_path, encoding="utf-8") as f:
        srt = f.read()
    os.unlink(raw_path)
    os.unlink(srt_path)
    with open(dst, "w", encoding="utf-8") as f:
        json.dump(dict(raw_lines=raw, srt=srt), f, ensure_ascii=False)


if __name__ == "__main__":
    main()
