"""Reading order, ROI / score filter and raw.txt format (video_subtitle_extractor_b200/rawtxt.py) against the outputs of
the reference's own `OcrRecogniser.predict`, `get_coordinates` and `extract_subtitles` on seeded inputs
(tests/golden/rawtxt_golden.json, written by tests/golden/make_rawtxt_golden.py where /root/reference exists)."""
import json
import os

import numpy as np

from video_subtitle_extractor_b200 import rawtxt

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "rawtxt_golden.json")


def _cases():
    with open(GOLDEN, encoding="utf-8") as f:
        return json.load(f)["cases"]


def test_reading_order_and_coordinates_match_reference_predict():
    cases = _cases()
    assert len(cases) >= 50 and sum(len(c["quads"]) for c in cases) > 100
    for c in cases:
        quads = [np.asarray(q, np.float32) for q in c["quads"]]
        rec = [(t, s) for t, s in c["rec"]]
        dt_box, res = rawtxt.order_like_predict(quads, rec)
        assert [[list(p) for p in b] for b in dt_box] == c["predict_boxes"]
        assert [[t, s] for t, s in res] == c["predict_res"]
        assert [list(x) for x in rawtxt.get_coordinates(dt_box)] == c["coordinates"]
    assert rawtxt.get_coordinates(np.zeros((0, 4, 2))) == []          # not a list -> [] (reference ocr.py:122)
    assert rawtxt.order_like_predict([], []) == ([], [])


def test_raw_txt_lines_match_reference_extract_subtitles():
    n_lines = n_dropped = 0
    for c in _cases():
        dt_box = [[tuple(p) for p in b] for b in c["predict_boxes"]]
        res = [(t, s) for t, s in c["predict_res"]]
        a, o = c["sub_area"], c["options"]
        area = (a["xmin"], a["xmax"], a["ymin"], a["ymax"]) if a else None
        got = rawtxt.frame_lines(c["frame_no"], dt_box, res, area, o["REC_CHAR_TYPE"], o["DROP_SCORE"], o["SUB_AREA_DEVIATION_RATE"])
        assert got == c["raw_lines"], (c["frame_no"], got, c["raw_lines"])
        n_lines += len(got)
        n_dropped += len(res) - len(got)
    assert n_lines > 50 and n_dropped > 20        # both branches of the filter are exercised


def test_y_round_and_overflow_rate():
    assert [rawtxt.y_round(y) for y in (590, 594, 595, 596, 600)] == [590, 590, 590, 600, 600]
    assert rawtxt.overflow_rate((0, 100, 0, 50), (10, 90, 10, 40)) == 0.0          # inside the area
    assert rawtxt.overflow_rate((0, 100, 0, 50), (200, 300, 0, 50)) is None        # no contact
    assert abs(rawtxt.overflow_rate((0, 100, 0, 50), (50, 150, 0, 50)) - 0.5) < 1e-12
    assert rawtxt.overflow_rate((0, 100, 0, 50), (100, 150, 0, 50)) == 0.5         # touching edge: kept by shapely as a line


def test_lines_from_engine_frame_result():
    from types import SimpleNamespace
    from video_subtitle_extractor_b200 import charset
    chars = charset.characters("en")
    # SURVEY.md Appendix E anchor (test_en.mp4 frame 300): "As far" = ids [18, 68, 96, 55, 50, 67]
    quad_a = np.array([[454, 642], [820, 649], [819, 689], [453, 682]], np.float32)
    quad_b = np.array([[979, 31], [1222, 31], [1222, 60], [979, 60]], np.float32)
    res = SimpleNamespace(quads=np.stack([quad_b, quad_a]), ids=[[18, 68], [18, 68, 96, 55, 50, 67]], rec_scores=np.array([0.99, 0.97], np.float32))
    lines = rawtxt.lines_from_frame_result(300, res, chars, sub_area=(64, 1216, 562, 713), rec_char_type="en", drop_score=0.75)
    assert lines == ["00000300\t(454, 819, 650, 682)\tAs far\n"]      # the title at the top of the frame is outside the default ROI
    assert len(rawtxt.lines_from_frame_result(300, res, chars)) == 2  # no ROI: everything is kept, top line first
