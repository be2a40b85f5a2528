"""Deterministic synthetic subtitle frames (BASELINE.json configs[1] and [4]).

Stands in for what the reference's frame reader hands to the predictor
(reference backend/tools/subtitle_ocr.py:164-208 -> ``predict(frame)``):
BGR uint8 HWC frames with 1-2 centred subtitle lines inside the default ROI
(reference backend/config.py:49, y in [0.78H, 0.99H]).

Recipe: smooth random background (1/32-resolution RGB field,
bilinearly upsampled) + N(0,4) noise, new background every 90 frames; white text
with a 3 px black stroke at 0.05*H px; a subtitle persists 45 frames, then 15 blank
frames.  The font is Pillow's embedded scalable default face so the generator is
self-contained on a machine without the reference tree.
"""
from __future__ import annotations

from typing import List, Tuple

import cv2
import numpy as np
from PIL import Image, ImageDraw, ImageFont

_LETTERS = "abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ"
_DIGITS = "0123456789"
_TAIL = [".", ",", "!", "?", "", "", "", ""]


class SynthStream:
    def __init__(self, height: int = 1080, width: int = 1920, seed: int = 20260117, blank_every: bool = True):
        self.h, self.w, self.seed = height, width, seed
        self.font_px = max(12, int(round(0.05 * height)))
        self.font = ImageFont.load_default(size=self.font_px)
        self.blank_every = blank_every
        self._bg_cache = (-1, None)
        self._txt_cache = (-1, None, None)

    # -- pieces ----------------------------------------------------------- #
    def _background(self, clip: int) -> np.ndarray:
        if self._bg_cache[0] == clip:
            return self._bg_cache[1]
        rng = np.random.default_rng([self.seed, 1, clip])
        small = rng.integers(0, 256, size=(max(2, self.h // 32), max(2, self.w // 32), 3), dtype=np.uint8)
        bg = cv2.resize(small, (self.w, self.h), interpolation=cv2.INTER_LINEAR)
        self._bg_cache = (clip, bg)
        return bg

    def _words(self, rng: np.random.Generator, n_words: int) -> str:
        words = []
        for _ in range(n_words):
            ln = int(rng.integers(2, 10))
            if rng.random() < 0.12:
                w = "".join(_DIGITS[int(i)] for i in rng.integers(0, 10, size=min(ln, 4)))
            else:
                w = "".join(_LETTERS[int(i)] for i in rng.integers(0, 26, size=ln))
                if rng.random() < 0.25:
                    w = w.capitalize()
            w += _TAIL[int(rng.integers(0, len(_TAIL)))]
            words.append(w)
        return " ".join(words)

    def _subtitle(self, sub: int) -> Tuple[np.ndarray, np.ndarray, List[str]]:
        """-> (alpha-premultiplied BGR overlay rows, alpha mask, lines) for the ROI band."""
        if self._txt_cache[0] == sub:
            return self._txt_cache[1], self._txt_cache[2]
        rng = np.random.default_rng([self.seed, 2, sub])
        n_lines = 1 if rng.random() < 0.5 else 2
        lines: List[str] = []
        max_w = int(self.w * 0.9)
        for _ in range(n_lines):
            n_words = int(rng.integers(4, 10))
            while True:
                s = self._words(rng, n_words)
                box = self.font.getbbox(s, stroke_width=3)
                if box[2] - box[0] <= max_w or n_words <= 2:
                    break
                n_words -= 1
            lines.append(s)
        img = Image.new("RGB", (self.w, self.h), (0, 0, 0))
        mask = Image.new("L", (self.w, self.h), 0)
        d, dm = ImageDraw.Draw(img), ImageDraw.Draw(mask)
        y_top, y_bot = 0.78 * self.h, 0.99 * self.h
        line_h = (y_bot - y_top) / 2.0
        for li, s in enumerate(lines):
            cy = y_top + line_h * (li + 0.5) if n_lines == 2 else y_top + line_h * 1.2
            d.text((self.w / 2, cy), s, font=self.font, fill=(255, 255, 255), stroke_width=3, stroke_fill=(0, 0, 0),
                   anchor="mm")
            dm.text((self.w / 2, cy), s, font=self.font, fill=255, stroke_width=3, stroke_fill=255, anchor="mm")
        out = (np.asarray(img)[:, :, ::-1].copy(), np.asarray(mask).copy())
        self._txt_cache = (sub, out, lines)
        return out, lines

    # -- public ----------------------------------------------------------- #
    def truth(self, i: int) -> List[str]:
        if self.blank_every and (i % 60) >= 45:
            return []
        return list(self._subtitle(i // 60)[1])

    def frame(self, i: int) -> np.ndarray:
        bg = self._background(i // 90)
        rng = np.random.default_rng([self.seed, 3, i])
        noise = rng.normal(0.0, 4.0, size=(self.h, self.w, 3)).astype(np.float32)
        fr = bg.astype(np.float32) + noise
        if not (self.blank_every and (i % 60) >= 45):
            (txt, mask), _ = self._subtitle(i // 60)
            a = (mask.astype(np.float32) / 255.0)[:, :, None]
            fr = fr * (1.0 - a) + txt.astype(np.float32) * a
        return np.clip(np.rint(fr), 0, 255).astype(np.uint8)

    def batch(self, start: int, n: int) -> List[np.ndarray]:
        return [self.frame(start + k) for k in range(n)]
