"""Golden vectors for the SRT writer: the reference's own ``SubtitleExtractor.generate_subtitle_file``
and ``_frame_to_timecode`` (backend/main.py:614-636, 731-766) run on seeded raw.txt files against the reference's sample
video test/test_en.mp4 (timecodes come from the decoder's CAP_PROP_POS_MSEC after a seek + read; recorded here per frame so
that the test needs no video).  Stubs as in make_dedup_golden.py.  Only runs where /root/reference exists."""
import json
import os
import sys
import tempfile
import types

import cv2
import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "srt_golden.json")
VIDEO = os.path.join(REF, "test", "test_en.mp4")


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
    finally:
        os.chdir(cwd)
    if "Main" not in m.tr:              # the stubbed config has no interface language: load the English strings
        m.tr.read(os.path.join(REF, "backend", "interface", "en.ini"), encoding="utf-8")
    cap = cv2.VideoCapture(VIDEO)
    fps, n_frames = cap.get(cv2.CAP_PROP_FPS), int(cap.get(cv2.CAP_PROP_FRAME_COUNT))
    cap.release()
    rng = np.random.default_rng(20260119)
    phrases = ["As far as we can go.", "Yami Sukehiro", "Let's get out of here!", "I don't know.", "Where are you going?", "OK"]
    ms_table = {}
    real_tc = m.SubtitleExtractor._frame_to_timecode

    def record_ms(frame_no):            # the same decoder calls as _frame_to_timecode, recorded for the test
        c = cv2.VideoCapture(VIDEO)
        c.set(cv2.CAP_PROP_POS_FRAMES, frame_no)
        ret, _ = c.read()
        ms_table[str(frame_no)] = float(c.get(cv2.CAP_PROP_POS_MSEC)) if ret else None
        c.release()

    cases = []
    for c in range(6):
        lines, frame = [], int(rng.integers(0, 40))
        for _ in range(int(rng.integers(3, 8))):
            text = str(rng.choice(phrases))
            for _ in range(int(rng.integers(1, 25))):
                lines.append(f"{str(frame).zfill(8)}\t(200, 800, 590, 630)\t{text}\n")
                frame += int(rng.choice([1, 1, 2, 3]))
            frame += int(rng.integers(0, 120))
        if c == 5:                       # frames past the end of the video: the decoder fails, the frame-count fallback is used
            lines.append(f"{str(n_frames + 500).zfill(8)}\t(200, 800, 590, 630)\tbeyond the end\n")
            lines.append(f"{str(n_frames + 700).zfill(8)}\t(200, 800, 590, 630)\tbeyond the end\n")
        with tempfile.NamedTemporaryFile("w", suffix=".txt", delete=False, encoding="utf-8") as f:
            f.writelines(lines)
            raw = f.name
        srt = raw + ".srt"
        fake = types.SimpleNamespace(raw_subtitle_path=raw, use_vsf=False, subtitle_output_path=srt, video_path=VIDEO, fps=fps,
                                     append_output=lambda *a, **k: None)
        fake._concat_content_with_same_frameno = lambda _f=fake: m.SubtitleExtractor._concat_content_with_same_frameno(_f)
        fake._remove_duplicate_subtitle = lambda _f=fake: m.SubtitleExtractor._remove_duplicate_subtitle(_f)

        def tc(frame_no, _f=fake):
            record_ms(frame_no)
            return real_tc(_f, frame_no)

        fake._frame_to_timecode = tc
        short = m.SubtitleExtractor.generate_subtitle_file(fake)
        with open(srt, encoding="utf-8") as f:
            text = f.read()
        os.unlink(raw)
        os.unlink(srt)
        cases.append(dict(lines=lines, srt=text, short_lines=short))
    with open(OUT, "w", encoding="utf-8") as f:
        json.dump(dict(generator="tests/golden/make_srt_golden.py", video="test/test_en.mp4", fps=fps, n_frames=n_frames,
                       reference_functions=["backend/main.py:614-636 generate_subtitle_file", "backend/main.py:731-766 _frame_to_timecode"],
                       pos_msec_after_seek_and_read=ms_table, cases=cases), f, ensure_ascii=False, indent=0)
    print(len(cases), "cases,", len(ms_table), "timecodes, fps", fps, "->", OUT)


if __name__ == "__main__":
    main()
