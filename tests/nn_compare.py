"""Shared helper: run a packed plan on the GPU engine and on the CPU plan interpreter, compare every value."""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from oracle.plan_interp import PlanInterpreter
from video_subtitle_extractor_b200 import plan as P


def to_bgrx(img_bgr: np.ndarray) -> np.ndarray:
    out = np.zeros(img_bgr.shape[:2] + (4,), np.uint8)
    out[:, :, :3] = img_bgr
    return out


def interp_values(plan: P.Plan, images_bgr: Sequence[np.ndarray], valid_w: Sequence[int]) -> Dict[int, np.ndarray]:
    """Per-image CPU runs (batch 1 each), concatenated pixel-major like the engine's ragged layout."""
    interp = PlanInterpreter(plan)
    per_value: Dict[int, List[np.ndarray]] = {}
    for img, vw in zip(images_bgr, valid_w):
        x = torch.from_numpy(np.ascontiguousarray(img))[None]
        _, env = interp.run(x, valid_w=[vw], keep_all=True)
        for vid, t in env.items():
            if vid == plan.input_vid:
                continue
            if t.dim() == 4:
                arr = t.permute(0, 2, 3, 1).reshape(-1, t.shape[1]).numpy()
            else:
                arr = t.reshape(1, -1).numpy()
            per_value.setdefault(vid, []).append(arr)
    return {vid: np.concatenate(parts, 0) for vid, parts in per_value.items()}


def compare_all(engine, which: int, plan: P.Plan, images_bgr: Sequence[np.ndarray], valid_w: Optional[Sequence[int]] = None,
                keep: Optional[dict] = None):
    """-> list of (step index, op name, out vid, max abs err, ref max abs) for every materialised step output.
    `keep`: optional dict; filled with {vid: (engine array, interpreter array)} for the plan's output values."""
    valid_w = list(valid_w) if valid_w is not None else [im.shape[1] for im in images_bgr]
    ref = interp_values(plan, images_bgr, valid_w)
    engine.debug_run_plan(which, [to_bgrx(im) for im in images_bgr], valid_w, keep_all=True)
    report = []
    for k, s in enumerate(plan.steps):
        if s.op == P.OP_COPY:
            continue
        got = engine.debug_get_value(which, s.out)
        if got is None or s.out not in ref:
            report.append((k, P.OP_NAMES[s.op], s.out, float("nan"), float("nan")))
            continue
        want = ref[s.out]
        if got.shape != want.shape:
            report.append((k, P.OP_NAMES[s.op], s.out, float("inf"), float(np.abs(want).max())))
            continue
        report.append((k, P.OP_NAMES[s.op], s.out, float(np.abs(got - want).max()), float(np.abs(want).max())))
        if keep is not None and s.out in plan.output_vids:
            keep[s.out] = (got, want)
    return report
