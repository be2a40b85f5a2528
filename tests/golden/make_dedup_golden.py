"""Golden vectors for the raw.txt post-processing that follows the OCR pass:

* ``SubtitleExtractor._concat_content_with_same_frameno`` (backend/main.py:820-864)
* ``SubtitleExtractor._remove_duplicate_subtitle``        (backend/main.py:774-818)

Runs the REFERENCE'S OWN methods (imported from /root/reference, unmodified, called on a stand-in ``self`` that carries the
three attributes they read) on seeded raw.txt files and records the rewritten file and the (start, end, text) list.
Module-level imports of the reference that are absent from this image are stubbed as in make_rawtxt_golden.py.
``Levenshtein.ratio`` is one of them: it is restated from its published definition (python-Levenshtein / rapidfuzz:
normalised indel similarity, 1 - (len(a) + len(b) - 2 LCS(a, b)) / (len(a) + len(b)), 1.0 for two empty strings) —
so the reference's control flow is pinned by these vectors, the similarity function by its definition.
"""
import json
import os
import sys
import tempfile
import types
from unittest.mock import MagicMock

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "dedup_golden.json")


def indel_ratio(a: str, b: str) -> float:
    if not a and not b:
        return 1.0
    prev = [0] * (len(b) + 1)
    for x in a:
        cur = [0]
        for j, y in enumerate(b, 1):
            cur.append(prev[j - 1] + 1 if x == y else max(prev[j], cur[j - 1]))
        prev = cur
    return 2.0 * prev[-1] / (len(a) + len(b))


def main():
    sys.path.insert(0, HERE)
    import make_rawtxt_golden as g
    g._stub_modules()
    lev = types.ModuleType("Levenshtein")
    lev.ratio = indel_ratio
    sys.modules["Levenshtein"] = lev
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
    assert m.config.thresholdTextSimilarity.value == 80
    rng = np.random.default_rng(20260118)
    phrases = ["As far as we can go.", "Yami Sukehiro", "Let's get out of here!", "字幕提取测试", "これはテストです", "Ｆｕｌｌ ｗｉｄｔｈ １２３",
               "I don't know.", "I don't knew.", "Where are you going?", "안녕하세요", "OK", "ok."]

    def noisy(t):
        if rng.random() < 0.35 and len(t) > 3:
            k = int(rng.integers(0, len(t)))
            t = t[:k] + str(rng.choice(list("abcxyz.,l1 "))) + t[k + 1:]
        if rng.random() < 0.15:
            t = t[:-1]
        return t

    cases = []
    for c in range(40):
        lines, frame = [], int(rng.integers(0, 50))
        for _ in range(int(rng.integers(1, 7))):
            text, second = str(rng.choice(phrases)), (str(rng.choice(phrases)) if rng.random() < 0.3 else None)
            for _ in range(int(rng.integers(1, 6))):
                y = int(rng.integers(590, 600))
                lines.append(f"{str(frame).zfill(8)}\t({int(rng.integers(100, 300))}, {int(rng.integers(700, 900))}, {y}, {y + 40})\t{noisy(text)}\n")
                if second and rng.random() < 0.8:
                    lines.append(f"{str(frame).zfill(8)}\t({int(rng.integers(100, 300))}, {int(rng.integers(700, 900))}, {y + 50}, {y + 90})\t{noisy(second)}\n")
                frame += int(rng.choice([1, 1, 1, 3, 8]))
            frame += int(rng.integers(0, 40))
        use_vsf = bool(c % 4 == 3)
        with tempfile.NamedTemporaryFile("w", suffix=".txt", delete=False, encoding="utf-8") as f:
            f.writelines(lines)
            path = f.name
        fake = types.SimpleNamespace(raw_subtitle_path=path, use_vsf=use_vsf)
        fake._concat_content_with_same_frameno = lambda _f=fake: m.SubtitleExtractor._concat_content_with_same_frameno(_f)
        unique = m.SubtitleExtractor._remove_duplicate_subtitle(fake)
        with open(path, encoding="utf-8") as f:
            rewritten = f.readlines()
        os.unlink(path)
        cases.append(dict(lines=lines, use_vsf=use_vsf, threshold=0.8, rewritten=rewritten, unique=[list(u) for u in unique]))
    with open(OUT, "w", encoding="utf-8") as f:
        json.dump(dict(generator="tests/golden/make_dedup_golden.py", reference_functions=[
            "backend/main.py:820-864 _concat_content_with_same_frameno", "backend/main.py:774-818 _remove_duplicate_subtitle"],
            cases=cases), f, ensure_ascii=False, indent=0)
    print(len(cases), "cases,", sum(len(c["unique"]) for c in cases), "subtitles from", sum(len(c["lines"]) for c in cases), "raw lines ->", OUT)


if __name__ == "__main__":
    main()
