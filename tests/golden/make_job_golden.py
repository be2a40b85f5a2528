"""Generates tests/golden/job_golden_<video>.json: raw.txt lines and .srt text written by the REFERENCE'S OWN glue
(OcrRecogniser.predict -> extract_subtitles -> _remove_duplicate_subtitle -> generate_subtitle_file, run untouched in its own
process by ref_glue_runner.py) from the graph-level oracle's predictor outputs over the whole fast-mode schedule of a sample
video (video_golden_<video>.json, make_video_golden.py).  Run HERE (needs /root/reference); the JSON is committed.
tests/test_gpu_jobs.py pipes the ENGINE's results through job.py (rawtxt.py / dedup.py) and compares with these files."""
import json
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from video_subtitle_extractor_b200 import charset  # noqa: E402
from video_subtitle_extractor_b200.job import default_sub_area  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = [("test_en", "en", 97), ("test_cn", "ch", 6625)]


def main():
    for name, lang, n_classes in CASES:
        with open(os.path.join(HERE, f"video_golden_{name}.json")) as f:
            g = json.load(f)
        chars = charset.characters(lang, None, n_classes)
        h, w = g["frames"][0]["shape"][:2]
        area = default_sub_area(h, w)
        frames = [dict(no=fr["no"], quads=[[[float(x), float(y)] for x, y in b] for b in fr["boxes"]],
                       rec=[[charset.ids_to_text(i, chars), s] for i, s in zip(fr["ids"], fr["rec_scores"])]) for fr in g["frames"]]
        job = dict(video=f"/root/reference/test/{g['video']}", fps=g["fps"],
                   area=dict(xmin=area[0], xmax=area[1], ymin=area[2], ymax=area[3]), frames=frames,
                   options=dict(REC_CHAR_TYPE=lang, DROP_SCORE=0.75, SUB_AREA_DEVIATION_RATE=0.0, DEBUG_OCR_LOSS=False))
        with tempfile.TemporaryDirectory() as tmp:
            with open(os.path.join(tmp, "in.json"), "w", encoding="utf-8") as f:
                json.dump(job, f, ensure_ascii=False)
            p = subprocess.run([sys.executable, os.path.join(HERE, "ref_glue_runner.py"), os.path.join(tmp, "in.json"),
                                os.path.join(tmp, "out.json")], capture_output=True, text=True, timeout=1800)
            assert p.returncode == 0, p.stderr[-2000:]
            with open(os.path.join(tmp, "out.json"), encoding="utf-8") as f:
                ref = json.load(f)
        out = dict(video=g["video"], lang=lang, area=list(area), drop_score=0.75, raw_lines=ref["raw_lines"], srt=ref["srt"])
        with open(os.path.join(HERE, f"job_golden_{name}.json"), "w", encoding="utf-8") as f:
            json.dump(out, f, ensure_ascii=False, indent=0)
        print(name, len(ref["raw_lines"]), "raw lines,", ref["srt"].count(" --> "), "subtitles")


if __name__ == "__main__":
    main()
