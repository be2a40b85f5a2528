"""raw.txt post-processing (video_subtitle_extractor_b200/dedup.py) against the outputs of the reference's own
`_concat_content_with_same_frameno` / `_remove_duplicate_subtitle` on seeded files (tests/golden/dedup_golden.json, written by
tests/golden/make_dedup_golden.py where /root/reference exists)."""
import json
import os

from video_subtitle_extractor_b200 import dedup

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "dedup_golden.json")


def test_ratio_is_the_normalised_indel_similarity():
    assert dedup.ratio("", "") == 1.0 and dedup.ratio("abc", "") == 0.0
    assert dedup.ratio("abc", "abc") == 1.0
    assert abs(dedup.ratio("kitten", "sitting") - 2 * 4 / 13) < 1e-12       # LCS "ittn"
    assert dedup.ratio("ab", "ba") == 0.5 == dedup.ratio("ba", "ab")


def test_concat_and_dedup_match_reference_methods():
    with open(GOLDEN, encoding="utf-8") as f:
        cases = json.load(f)["cases"]
    assert len(cases) >= 40
    merged = subtitles = 0
    for c in cases:
        got_lines = dedup.concat_same_frame(c["lines"])
        assert got_lines == c["rewritten"]
        merged += len(c["lines"]) - len(got_lines)
        got = dedup.remove_duplicates(c["lines"], c["threshold"], c["use_vsf"])
        assert [list(u) for u in got] == c["unique"]
        subtitles += len(got)
    assert merged > 30 and subtitles > 100


def test_srt_writer_matches_reference_generate_subtitle_file():
    with open(os.path.join(os.path.dirname(GOLDEN), "srt_golden.json"), encoding="utf-8") as f:
        g = json.load(f)
    table = g["pos_msec_after_seek_and_read"]
    n = 0
    for c in g["cases"]:
        subs = dedup.remove_duplicates(c["lines"], 0.8, use_vsf=False)
        text, short = dedup.srt_text(subs, g["fps"], lambda frame_no: table[str(frame_no)])
        assert text == c["srt"]
        assert short == c["short_lines"]
        n += len(subs)
    assert n > 25 and any(v is None for v in table.values())      # the decoder-failure fallback is exercised too
    assert dedup.timecode_from_msec(3723004.9) == "01:02:03,004"
    assert dedup.timecode_from_frame(4597, 29.97002997002997) == "00:02:33,011"
