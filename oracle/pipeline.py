"""ORACLE (test infrastructure, never shipped on the product path).

CPU restatement of the per-frame det+rec path exactly as the reference drives it:
one frame per call, det batch 1, rec in aspect-sorted batches of ``rec_batch_num``.

* ``OraclePipeline.detect``  ~ reference backend/tools/subtitle_detect.py:24-26
* ``OraclePipeline.ocr``     ~ reference backend/tools/ocr.py:27 (``PaddleOCR.__call__`` = TextSystem)

NN arithmetic: oracle/graph_interp.py (shipped graphs on torch-CPU fp32);
host logic: oracle/hostlogic.py.  Parity unpinned by the reference.
"""
from __future__ import annotations

import copy
import time
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np

from video_subtitle_extractor_b200.loader import load_model

from . import hostlogic as hl
from .graph_interp import GraphInterpreter


@dataclass
class OracleResult:
    boxes: List[np.ndarray]              # float32 [4,2] each, TextSystem order
    ids: List[List[int]]                 # CTC class ids per box
    scores: List[float]                  # mean kept max-prob
    rec_widths: List[int] = field(default_factory=list)   # padded imgW each crop was run at
    det_scores: List[float] = field(default_factory=list)


class _PlanNet:
    """Adapter: run a packed plan (the committed .vsep files) on the CPU plan interpreter with the
    GraphInterpreter calling convention (float NCHW in, reference-layout outputs out).  Used where the reference
    tree is absent (GPU box); tests/test_plan_cpu.py proves plan == graph where it is present."""

    def __init__(self, blob: bytes, is_rec: bool):
        from video_subtitle_extractor_b200 import plan as P
        from .plan_interp import PlanInterpreter
        self.interp = PlanInterpreter(P.deserialize(blob))
        self.is_rec = is_rec

    def run(self, x):
        import torch
        if isinstance(x, np.ndarray):
            x = torch.from_numpy(np.ascontiguousarray(x))
        outs = self.interp.run(x.float())
        if self.is_rec:  # [B, C, 1, T] -> [B, T, C]
            outs = [o[:, :, 0, :].permute(0, 2, 1).contiguous() for o in outs]
        return outs


class OraclePipeline:
    @classmethod
    def from_plans(cls, det_blob: bytes, rec_blob: Optional[bytes] = None, **kw) -> "OraclePipeline":
        self = cls.__new__(cls)
        self.det = _PlanNet(det_blob, False)
        self.rec = _PlanNet(rec_blob, True) if rec_blob else None
        self.rec_batch_num = kw.get("rec_batch_num", 6)
        shape = kw.get("rec_image_shape", (3, 48, 320))
        self.rec_h, self.rec_w = shape[1], shape[2]
        self.det_fetch = 0
        self.limit_side_len = kw.get("limit_side_len", 960)
        return self

    def __init__(self, det_model_dir: str, rec_model_dir: Optional[str] = None, rec_batch_num: int = 6,
                 rec_image_shape: Tuple[int, int, int] = (3, 48, 320), det_fetch: int = 0,
                 limit_side_len: int = 960):
        self.det = GraphInterpreter(load_model(det_model_dir))
        self.rec = GraphInterpreter(load_model(rec_model_dir)) if rec_model_dir else None
        self.rec_batch_num = rec_batch_num
        self.rec_h, self.rec_w = rec_image_shape[1], rec_image_shape[2]
        self.det_fetch = det_fetch
        self.limit_side_len = limit_side_len

    # -- detection -------------------------------------------------------- #
    def det_prob_map(self, img: np.ndarray):
        x, shape, _ = hl.det_preprocess(img, self.limit_side_len)
        outs = self.det.run(x)
        return outs[self.det_fetch][0, 0].numpy(), shape

    def detect(self, img: np.ndarray, return_scores: bool = False):
        """-> float32 [N,4,2] quads in frame pixels (TextDetector.__call__)."""
        pred, shape = self.det_prob_map(img)
        boxes, scores = hl.db_postprocess(pred, shape)
        # filter_tag_det_res may drop boxes; keep scores aligned
        kept_boxes, kept_scores = [], []
        for b, s in zip(boxes, scores):
            fb = hl.filter_tag_det_res(b[None], img.shape)
            if len(fb):
                kept_boxes.append(fb[0])
                kept_scores.append(s)
        out = np.array(kept_boxes, dtype=np.float32).reshape(-1, 4, 2)
        if return_scores:
            return out, kept_scores
        return out

    # -- recognition ------------------------------------------------------ #
    def recognise(self, crops: Sequence[np.ndarray]):
        """-> [(ids, score)], padded widths; upstream batching (aspect sort, <=6, per-batch imgW)."""
        n = len(crops)
        results: List[Tuple[List[int], float]] = [([], 0.0)] * n
        widths = [0] * n
        ratios = [c.shape[1] / float(c.shape[0]) for c in crops]
        for idxs, img_w in hl.rec_batches(ratios, self.rec_batch_num, self.rec_h, self.rec_w):
            batch = np.stack([hl.resize_norm_img(crops[i], img_w, self.rec_h) for i in idxs])
            probs = self.rec.run(batch)[0].numpy()
            for k, i in enumerate(idxs):
                ids, score, _ = hl.ctc_decode_ids(probs[k])
                results[i] = (ids, score)
                widths[i] = img_w
        return results, widths

    def ocr(self, img: np.ndarray) -> OracleResult:
        """TextSystem.__call__(img, cls=False) with drop_score=0 (reference ocr.py:105)."""
        ori = img.copy()
        dt_boxes, det_scores = self.detect(img, return_scores=True)
        if len(dt_boxes) == 0:
            return OracleResult([], [], [])
        order = hl.sorted_boxes(dt_boxes)
        # keep det scores aligned with the sorted order
        score_of = {b.tobytes(): s for b, s in zip(dt_boxes, det_scores)}
        crops = [hl.get_rotate_crop_image(ori, copy.deepcopy(b)) for b in order]
        rec, widths = self.recognise(crops)
        return OracleResult(boxes=[np.array(b, dtype=np.float32) for b in order],
                            ids=[r[0] for r in rec], scores=[r[1] for r in rec], rec_widths=widths,
                            det_scores=[score_of.get(b.tobytes(), 0.0) for b in order])
