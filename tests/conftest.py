import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _has_gpu() -> bool:
    try:
        from video_subtitle_extractor_b200 import engine
        return engine.device_count() > 0
    except Exception:
        return False


HAS_GPU = None


def pytest_collection_modifyitems(config, items):
    global HAS_GPU
    markexpr = config.getoption("-m") or ""
    for item in items:
        if "gpu" in item.keywords:
            if HAS_GPU is None:
                HAS_GPU = _has_gpu()
            if not HAS_GPU and "gpu" not in markexpr.replace("not gpu", ""):
                item.add_marker(pytest.mark.skip(reason="no CUDA device"))
