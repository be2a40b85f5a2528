"""Golden vectors for the fast-mode frame schedule and the half-frame crop: the reference's own
``SubtitleExtractor.extract_frame_by_fps`` (backend/main.py:228-251) on a fake capture of n frames and ``frame_preprocess``
(backend/tools/subtitle_ocr.py:270-289) on arrays whose rows are numbered.  Stubs as in make_dedup_golden.py."""
import json
import os
import sys
import types

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "frames_golden.json")


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
        from backend.tools import subtitle_ocr as so
        from backend.tools.constant import SubtitleArea as AreaKind
    finally:
        os.chdir(cwd)
    schedule = []
    for n_frames in (0, 1, 2, 7, 29, 30, 31, 100, 3597):
        for fps in (23.976023976023978, 25.0, 29.97002997002997, 60.0, 2.0):
            for freq in (1, 3, 10, 60):
                m.config.extractFrequency.value = freq
                state, tasks = dict(pos=0), []

                class Cap:
                    def isOpened(self):
                        return True

                    def read(self):
                        if state["pos"] >= n_frames:
                            return False, None
                        state["pos"] += 1
                        return True, state["pos"]

                    def release(self):
                        pass

                fake = types.SimpleNamespace(video_cap=Cap(), frame_count=n_frames, fps=fps, update_progress=lambda **k: None,
                                             subtitle_ocr_task_queue=types.SimpleNamespace(put=lambda t: tasks.append(t)))
                m.SubtitleExtractor.extract_frame_by_fps(fake)
                schedule.append(dict(n_frames=n_frames, fps=fps, extract_frequency=freq, frames=[t[1] for t in tasks]))
    m.config.extractFrequency.value = 3
    crops = []
    for h in (1, 2, 7, 720, 1080, 1081):
        rows = np.arange(h, dtype=np.int32).reshape(h, 1, 1).repeat(3, axis=2)
        for kind, name in ((AreaKind.LOWER_PART, "lower"), (AreaKind.UPPER_PART, "upper"), (AreaKind.UNKNOWN, "unknown"), (None, None)):
            out = so.frame_preprocess(kind, rows) if kind is not None else rows
            crops.append(dict(h=h, kind=name, first_row=int(out[0, 0, 0]) if len(out) else None, n_rows=int(out.shape[0])))
    with open(OUT, "w", encoding="utf-8") as f:
        json.dump(dict(generator="tests/golden/make_frames_golden.py", reference_functions=[
            "backend/main.py:228-251 extract_frame_by_fps", "backend/tools/subtitle_ocr.py:270-289 frame_preprocess"],
            schedule=schedule, crops=crops), f, indent=0)
    print(len(schedule), "schedules,", len(crops), "crops ->", OUT)


if __name__ == "__main__":
    main()
