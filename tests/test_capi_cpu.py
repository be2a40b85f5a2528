"""The C-ABI library loads on a machine without a GPU, exports every symbol include/vse_b200.h declares, and fails
loudly (no CPU fallback) when asked to compute without a device."""
import ctypes as C
import os
import re

import pytest

from video_subtitle_extractor_b200 import engine as E

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "vse_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vse_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported():
    lib = E.load_library()
    names = declared_symbols()
    assert len(names) >= 14
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(E.EXPORTS) == names
    assert lib.vse_abi_version() == 1


def test_default_config_mirrors_reference_knobs():
    lib = E.load_library()
    cfg = E.VseConfig()
    lib.vse_default_config(C.byref(cfg))
    # reference backend/tools/ocr.py:91-113 + upstream defaults
    assert (cfg.det_limit_side_len, cfg.rec_image_h, cfg.rec_image_w, cfg.rec_batch_num) == (960, 48, 320, 6)
    assert abs(cfg.det_thresh - 0.3) < 1e-7 and abs(cfg.det_box_thresh - 0.6) < 1e-7 and abs(cfg.det_unclip_ratio - 1.5) < 1e-7
    assert cfg.det_max_candidates == 1000
    # the default mode is the one held to the parity bar (tests/test_gpu_real_video.py), not the faster fp16 storage
    assert cfg.precision == E.PRECISION_FP32_TC == E.bench_mode()["precision"] and cfg.flags == 0


def test_no_device_means_no_engine():
    if E.device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(RuntimeError, match="no CUDA device|CPU fallback"):
        E.Engine()


def test_missing_library_is_an_error(tmp_path, monkeypatch):
    monkeypatch.setattr(E, "_lib", None)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        E.load_library(str(tmp_path / "libvse_b200.so"))
