"""``utility.parse_args()`` stand-in: an attribute bag with the upstream defaults the reference relies on
(reference backend/tools/subtitle_detect.py:14-21 assigns det_algorithm / det_model_dir / use_gpu / use_onnx on it)."""
from types import SimpleNamespace


def parse_args():
    return SimpleNamespace(det_algorithm="DB", det_model_dir=None, det_limit_side_len=960, det_limit_type="max",
                           det_db_thresh=0.3, det_db_box_thresh=0.6, det_db_unclip_ratio=1.5, use_dilation=False,
                           det_db_score_mode="fast", det_box_type="quad", use_gpu=True, gpu_id=0, use_onnx=False,
                           onnx_providers=None, benchmark=False, show_log=False)
