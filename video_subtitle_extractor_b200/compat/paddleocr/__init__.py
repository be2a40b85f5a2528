"""Stand-in for ``paddleocr`` that forwards to the B200 engine.

Put ``video_subtitle_extractor_b200/compat`` FIRST on ``sys.path`` and the reference's untouched
``from paddleocr import PaddleOCR`` (backend/tools/ocr.py:4) resolves here.  Contract honoured:
``PaddleOCR(**kwargs)(image, cls=False) -> (boxes, rec_res, time_dict)`` with ``boxes`` a *list* of float32 [4,2]
(clockwise from top-left, frame pixels), ``rec_res`` a *list* of ``(str, float)``, empty frame -> ``([], [], {...})``;
unknown kwargs are accepted and ignored; results with ``score < drop_score`` are dropped (0 in the reference, ocr.py:105).
"""
from __future__ import annotations

import os
import time
from typing import Optional

import numpy as np

from video_subtitle_extractor_b200 import charset, engine as _E, weights as _W

__version__ = "2.10.0+vse_b200"


def _plan_for(model_dir: str) -> bytes:
    model_dir = os.path.normpath(model_dir)
    name = "/".join(model_dir.replace("\\", "/").split("/")[-2:])
    return _W.load_plan_blob(name, models_root=os.path.dirname(os.path.dirname(model_dir)))


class PaddleOCR:
    def __init__(self, det_model_dir: Optional[str] = None, rec_model_dir: Optional[str] = None, rec_batch_num: int = 6,
                 rec_image_shape: str = "3,48,320", lang: str = "ch", drop_score: float = 0.5, det: bool = True,
                 rec_char_dict_path: Optional[str] = None, det_limit_side_len: int = 960, det_db_thresh: float = 0.3,
                 det_db_box_thresh: float = 0.6, det_db_unclip_ratio: float = 1.5, gpu_id: int = 0, **ignored):
        shape = [int(v) for v in str(rec_image_shape).split(",")]
        self.drop_score = drop_score
        # one mode for every shipped model: fp32 activations, tensor-core products on fp16 hi/lo splits scaled per layer
        # (engine.bench_mode(): the mode the parity tests hold to the reference's results; the accurate-mode server detector,
        # paddle_model_config.py:60,70, exceeds the fp16 range and runs through its calibrated per-layer scales)
        self.engine = _E.Engine(device=gpu_id, rec_image_h=shape[1], rec_image_w=shape[2], rec_batch_num=rec_batch_num,
                                **_E.bench_mode(),
                                det_limit_side_len=det_limit_side_len, det_thresh=det_db_thresh,
                                det_box_thresh=det_db_box_thresh, det_unclip_ratio=det_db_unclip_ratio)
        self.engine.load_plan(_E.PLAN_DET, _plan_for(det_model_dir), det_model_dir)
        self.has_rec = rec_model_dir is not None
        if self.has_rec:
            blob = _plan_for(rec_model_dir)
            self.engine.load_plan(_E.PLAN_REC, blob, rec_model_dir)
            from video_subtitle_extractor_b200 import plan as _P
            p = _P.deserialize(blob)
            self.chars = charset.characters(lang, rec_char_dict_path, n_classes=p.values[p.output_vids[0]].channels)

    def __call__(self, img, cls: bool = False, **kw):
        t0 = time.time()
        frame = np.ascontiguousarray(img)
        if frame.ndim == 2:
            frame = np.stack([frame] * 3, axis=-1)
        r = self.engine.run([frame], det_only=not self.has_rec)[0]
        boxes, rec_res = [], []
        for k in range(len(r.quads)):
            score = float(r.rec_scores[k]) if self.has_rec else 0.0
            if self.has_rec and score < self.drop_score:
                continue
            boxes.append(r.quads[k].copy())
            rec_res.append((charset.ids_to_text(r.ids[k], self.chars) if self.has_rec else "", score))
        return boxes, rec_res, {"all": time.time() - t0}

    # paddleocr's convenience entry point
    def ocr(self, img, det=True, rec=True, cls=False, **kw):
        boxes, rec_res, _ = self(img, cls=cls)
        return [[[b.tolist(), r] for b, r in zip(boxes, rec_res)]]
