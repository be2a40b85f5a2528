"""``TextDetector`` stand-in (reference backend/tools/subtitle_detect.py:12,22,25): ``TextDetector(args)(img)`` returns
``(np.float32[N,4,2], elapse_seconds)`` — an ndarray (callers use ``.tolist()`` / ``len()``), contour order."""
import time

import numpy as np

from video_subtitle_extractor_b200 import engine as _E
from video_subtitle_extractor_b200 import weights as _W
from paddleocr import _plan_for


class TextDetector:
    def __init__(self, args):
        self.engine = _E.Engine(device=getattr(args, "gpu_id", 0), det_limit_side_len=args.det_limit_side_len, **_E.bench_mode(),
                                det_thresh=args.det_db_thresh, det_box_thresh=args.det_db_box_thresh,
                                det_unclip_ratio=args.det_db_unclip_ratio)
        self.engine.load_plan(_E.PLAN_DET, _plan_for(args.det_model_dir), args.det_model_dir)

    def __call__(self, img):
        t0 = time.time()
        r = self.engine.run([np.ascontiguousarray(img)], det_only=True)[0]
        return r.quads.astype(np.float32).reshape(-1, 4, 2), time.time() - t0
