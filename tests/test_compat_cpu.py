"""The drop-in shim resolves the reference's three import sites without a GPU."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COMPAT = os.path.join(ROOT, "video_subtitle_extractor_b200", "compat")


@pytest.fixture()
def compat_path(monkeypatch):
    monkeypatch.syspath_prepend(COMPAT)
    for m in [k for k in sys.modules if k == "paddle" or k.startswith("paddleocr")]:
        monkeypatch.delitem(sys.modules, m)
    yield


def test_reference_import_sites_resolve(compat_path):
    from paddleocr import PaddleOCR                                   # backend/tools/ocr.py:4
    from paddleocr.tools.infer import utility                          # backend/tools/subtitle_detect.py:11
    from paddleocr.tools.infer.predict_det import TextDetector         # backend/tools/subtitle_detect.py:12
    import paddle                                                      # backend/tools/hardware_accelerator.py:2
    args = utility.parse_args()
    args.det_algorithm, args.det_model_dir, args.use_gpu, args.use_onnx = "DB", "x/V4/ch_det_fast", True, False
    assert (args.det_limit_side_len, args.det_db_thresh, args.det_db_box_thresh, args.det_db_unclip_ratio) == (960, 0.3, 0.6, 1.5)
    assert isinstance(paddle.is_compiled_with_cuda(), bool)
    assert len(paddle.static.cuda_places()) == (1 if paddle.is_compiled_with_cuda() else 0) or paddle.is_compiled_with_cuda()
    assert callable(PaddleOCR) and callable(TextDetector)


def test_shim_has_no_cpu_fallback(compat_path):
    from paddleocr import PaddleOCR
    from video_subtitle_extractor_b200 import engine as E
    if E.device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(RuntimeError):
        PaddleOCR(det_model_dir="/x/V4/ch_det_fast", rec_model_dir="/x/V4/en_rec_fast", lang="en", drop_score=0)


def test_charset():
    from video_subtitle_extractor_b200 import charset
    chars = charset.characters("en")
    assert charset.ids_to_text([18, 68, 96, 55, 50, 67], chars) == "As far"
    assert len(charset.characters("ch", n_classes=6625)) == 6625
