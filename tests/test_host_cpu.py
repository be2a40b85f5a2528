"""Host-side logic restated from paddleocr (oracle/hostlogic.py) and the synthetic frame generator."""
import math

import numpy as np

from oracle import hostlogic as hl
from video_subtitle_extractor_b200.synth import SynthStream


def test_det_resize_shape_matches_survey_table():
    # SURVEY.md §8a row a8
    assert hl.det_resize_shape(1080, 1920) == (544, 960)
    assert hl.det_resize_shape(720, 1280) == (544, 960)
    assert hl.det_resize_shape(2160, 3840) == (544, 960)
    assert hl.det_resize_shape(1080, 1440) == (704, 960)      # 22.5 -> 22 (banker's rounding)
    assert hl.det_resize_shape(886, 1920) == (448, 960)
    assert hl.det_resize_shape(20, 50) == (32, 64)


def test_rec_batches_follow_upstream_rules():
    ratios = [10.0, 3.0, 25.0, 7.0, 6.9, 12.0, 4.0]
    batches = list(hl.rec_batches(ratios, 6, 48, 320))
    assert [sorted(b[0]) for b in batches] == [[0, 1, 3, 4, 5, 6], [2]]
    assert batches[0][1] == int(48 * 12.0) and batches[1][1] == int(48 * 25.0)
    assert list(hl.rec_batches([2.0], 6, 48, 320))[0][1] == 320   # never narrower than rec_image_shape


def test_ctc_decode_collapses_and_drops_blank():
    p = np.zeros((7, 5), np.float32)
    for t, (c, v) in enumerate([(0, .9), (2, .8), (2, .7), (0, .6), (2, .5), (3, .4), (3, .3)]):
        p[t, c] = v
    ids, score, kept = hl.ctc_decode_ids(p)
    assert ids == [2, 2, 3]
    assert abs(score - np.mean([.8, .5, .4])) < 1e-6
    assert hl.ctc_decode_ids(np.eye(5, dtype=np.float32)[[0, 0, 0]]) == ([], 0.0, [])


def test_en_dictionary():
    assert len(hl.EN_CHARACTERS) == 97 and hl.EN_CHARACTERS[0] == "blank" and hl.EN_CHARACTERS[-1] == " "
    assert hl.ids_to_text([18, 68, 96, 55], hl.EN_CHARACTERS) == "As f"


def test_sorted_boxes_reading_order():
    def box(x, y):
        return np.array([[x, y], [x + 50, y], [x + 50, y + 20], [x, y + 20]], np.float32)
    out = hl.sorted_boxes(np.array([box(300, 104), box(10, 100), box(200, 300), box(100, 109)]))
    assert [tuple(b[0]) for b in out] == [(10, 100), (100, 109), (300, 104), (200, 300)]


def test_clipper_round_offset_of_a_rectangle():
    box = np.array([[10, 20], [110, 20], [110, 50], [10, 50]], np.float32)
    d = 9.0
    pts = hl.clipper_offset_round(box, d)
    assert pts[:, 0].min() == 1 and pts[:, 0].max() == 119 and pts[:, 1].min() == 11 and pts[:, 1].max() == 59
    # shoelace area of the offset polygon ~ rectangle grown by d with round corners
    x, y = pts[:, 0].astype(float), pts[:, 1].astype(float)
    area = abs(np.dot(x, np.roll(y, -1)) - np.dot(y, np.roll(x, -1))) / 2
    want = 100 * 30 + 2 * d * 130 + math.pi * d * d
    assert abs(area - want) / want < 0.01


def test_synthetic_stream_is_deterministic_and_labelled():
    a, b = SynthStream(270, 480), SynthStream(270, 480)
    assert np.array_equal(a.frame(7), b.frame(7))
    assert a.truth(50) == [] and 1 <= len(a.truth(0)) <= 2
    f = a.frame(0)
    assert f.shape == (270, 480, 3) and f.dtype == np.uint8
    band = f[int(0.78 * 270):, :, :]
    assert (band == 255).mean() > 0.005          # white glyphs inside the default ROI band
