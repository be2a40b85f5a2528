"""``paddle`` stand-in for the reference's capability probe (backend/tools/hardware_accelerator.py:2,26-32):
reports CUDA when the B200 engine sees a device, so `accurate` mode takes the per-frame det path (backend/main.py:140)
and `auto` picks the server models (backend/tools/paddle_model_config.py:59-65)."""
from types import SimpleNamespace

from video_subtitle_extractor_b200 import engine as _E

__version__ = "3.0.0+vse_b200"


def is_compiled_with_cuda() -> bool:
    return _E.device_count() > 0


def _cuda_places():
    return [f"Place(gpu:{i})" for i in range(_E.device_count())]


static = SimpleNamespace(cuda_places=_cuda_places)
device = SimpleNamespace(cuda=SimpleNamespace(device_count=_E.device_count))
