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
