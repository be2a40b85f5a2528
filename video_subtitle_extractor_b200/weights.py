"""Packed engine plans (.vsep) for the reference's shipped models.

``build_plan`` converts ``backend/models/<ver>/<name>`` of the reference into the binary the C++ runtime
loads (plan.py); ``load_plan_blob`` returns a previously packed plan.  The two models of the headline
benchmark (BASELINE.json configs[1]: V4/ch_det_fast + V4/en_rec_fast) are committed under ``weights/`` so
that tests, smoke() and bench.py run on a machine without the reference tree; the others are packed by
``__graft_entry__.build()`` wherever /root/reference exists.
"""
from __future__ import annotations

import os
from typing import Dict, Optional

from . import plan as P
from .loader import load_model

WEIGHTS_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "weights")
DEFAULT_MODELS = ["V4/ch_det_fast", "V4/en_rec_fast"]
EXTRA_MODELS = ["V4/ch_rec_fast", "V3/japan_rec_fast", "V3/korean_rec_fast", "V4/ch_det", "V4/ch_rec", "V2/ch_rec"]


def _path(name: str) -> str:
    return os.path.join(WEIGHTS_DIR, name.replace("/", "__") + ".vsep")


# models whose activations exceed the fp16 range: the engine must be created with fp32 activations
FP32_ONLY = {"V4/ch_det"}


def needs_fp32(name: str) -> bool:
    return "/".join(name.replace("\\", "/").rstrip("/").split("/")[-2:]) in FP32_ONLY


def is_det(name: str) -> bool:
    return "_det" in name


def build_plan(models_root: str, name: str, save: bool = True) -> bytes:
    model = load_model(os.path.join(models_root, name))
    scale, shift = P.DET_NORM if is_det(name) else P.REC_NORM
    plan = P.compile_model(model, name=name, norm_scale=scale, norm_shift=shift, fetch_cols=[0])
    blob = plan.serialize()
    if save:
        os.makedirs(WEIGHTS_DIR, exist_ok=True)
        tmp = _path(name) + ".tmp"
        with open(tmp, "wb") as f:
            f.write(blob)
        os.replace(tmp, _path(name))
    return blob


def build_default_plans(models_root: str, extra: bool = True) -> None:
    for name in DEFAULT_MODELS + (EXTRA_MODELS if extra else []):
        src = os.path.join(models_root, name, "inference.pdmodel")
        if not os.path.exists(src):
            continue
        out = _path(name)
        if os.path.exists(out) and os.path.getmtime(out) >= max(os.path.getmtime(src), os.path.getmtime(P.__file__)):
            continue
        build_plan(models_root, name)


def have_plan(name: str) -> bool:
    return os.path.exists(_path(name))


def load_plan_blob(name: str, models_root: Optional[str] = None) -> bytes:
    path = _path(name)
    if os.path.exists(path):
        with open(path, "rb") as f:
            return f.read()
    if models_root and os.path.exists(os.path.join(models_root, name)):
        return build_plan(models_root, name)
    raise FileNotFoundError(f"no packed plan for {name} ({path}); run __graft_entry__.build() where the reference models exist")
