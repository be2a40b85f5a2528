"""ORACLE (test infrastructure, never shipped on the product path).

CPU restatement of the host-side logic that ``paddleocr~=2.10.0``
(reference requirements.txt:16) runs around the two predictor calls.  paddleocr is a
pip dependency absent from /root/reference, so its published algorithm is restated
here with numpy + cv2 (cv2 is the reference's own dependency, requirements.txt:1);
pyclipper/shapely are absent too, so the Clipper polygon offset is restated in
``clipper_offset_round``.  Call sites this follows:

* reference backend/tools/subtitle_detect.py:24-26   TextDetector.__call__        -> ``det_preprocess`` / ``db_postprocess`` / ``filter_tag_det_res``
* reference backend/tools/ocr.py:27                  PaddleOCR.__call__ (TextSystem) -> ``sorted_boxes`` / ``get_rotate_crop_image``
* reference backend/tools/ocr.py:97-99,108           rec_algorithm/rec_batch_num/rec_image_shape -> ``rec_batches`` / ``resize_norm_img``
* CTC greedy decode (CTCLabelDecode)                                              -> ``ctc_decode``

Upstream behaviours taken from memory are listed in SURVEY.md Appendix D.8.
Parity unpinned by the reference (no tests, no golden vectors: SURVEY.md §4).
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence, Tuple

import cv2
import numpy as np

DET_MEAN = np.array([0.485, 0.456, 0.406], dtype=np.float32).reshape(1, 1, 3)
DET_STD = np.array([0.229, 0.224, 0.225], dtype=np.float32).reshape(1, 1, 3)

# en_dict.txt, reconstructed: 0x30-0x7E, 0x21-0x2F, 0x20
EN_DICT = "".join(chr(c) for c in range(0x30, 0x7F)) + "".join(chr(c) for c in range(0x21, 0x30)) + " "


def character_list(dict_chars: Optional[Sequence[str]], use_space_char: bool = True) -> List[str]:
    """CTCLabelDecode character table: ['blank'] + dict (+ ' ')."""
    chars = list(dict_chars) if dict_chars is not None else list("0123456789abcdefghijklmnopqrstuvwxyz")
    if use_space_char and dict_chars is not None:
        chars.append(" ")
    return ["blank"] + chars


EN_CHARACTERS = character_list(list(EN_DICT[:-1]) + [" "], use_space_char=True)  # 1 + 95 + 1 = 97


# --------------------------------------------------------------------------- #
# detection pre-process (DetResizeForTest + NormalizeImage + ToCHWImage)
# --------------------------------------------------------------------------- #

def det_resize_shape(h: int, w: int, limit_side_len: int = 960) -> Tuple[int, int]:
    """limit_type='max': scale so max side <= limit, then round each side to a multiple of 32."""
    if max(h, w) > limit_side_len:
        ratio = float(limit_side_len) / h if h > w else float(limit_side_len) / w
    else:
        ratio = 1.0
    rh = int(h * ratio)
    rw = int(w * ratio)
    rh = max(int(round(rh / 32) * 32), 32)  # Python round = half-to-even
    rw = max(int(round(rw / 32) * 32), 32)
    return rh, rw


def det_preprocess(img: np.ndarray, limit_side_len: int = 960):
    """BGR uint8 HWC -> (float32 [1,3,rh,rw], shape=[src_h, src_w, ratio_h, ratio_w])."""
    src_h, src_w = img.shape[:2]
    if src_h + src_w < 64:
        ph, pw = max(32, src_h), max(32, src_w)
        pad = np.zeros((ph, pw, img.shape[2]), dtype=np.uint8)
        pad[:src_h, :src_w] = img
        img = pad
    h, w = img.shape[:2]
    rh, rw = det_resize_shape(h, w, limit_side_len)
    resized = cv2.resize(img, (rw, rh))
    ratio_h = rh / float(h)
    ratio_w = rw / float(w)
    x = (resized.astype("float32") * np.float32(1.0 / 255.0) - DET_MEAN) / DET_STD
    x = x.transpose((2, 0, 1))[None]
    return np.ascontiguousarray(x, dtype=np.float32), np.array([src_h, src_w, ratio_h, ratio_w]), resized


# --------------------------------------------------------------------------- #
# Clipper polygon offset (pyclipper.PyclipperOffset, JT_ROUND, ET_CLOSEDPOLYGON)
# --------------------------------------------------------------------------- #

def _clipper_round(v: float) -> int:
    return int(v - 0.5) if v < 0 else int(v + 0.5)


def clipper_offset_round(path_float: np.ndarray, delta: float, arc_tolerance: float = 0.25) -> np.ndarray:
    """ClipperLib 6.4.2 ``ClipperOffset::DoOffset`` for one closed polygon with round joins.

    ``path_float`` [N,2] float coordinates; pyclipper truncates them to integers on AddPath.
    Returns the integer offset polygon vertices [M,2] (before Clipper's final union, which
    for the convex quads used here keeps the same outer boundary).
    """
    pts = [(int(p[0]), int(p[1])) for p in path_float]  # C-cast truncation
    # strip duplicates (ClipperOffset::AddPath)
    high = len(pts) - 1
    while high > 0 and pts[0] == pts[high]:
        high -= 1
    contour = [pts[0]]
    for i in range(1, high + 1):
        if contour[-1] != pts[i]:
            contour.append(pts[i])
    if len(contour) < 3:
        return np.zeros((0, 2), dtype=np.int64)
    # FixOrientations: Clipper area > 0 <=> orientation true; reverse otherwise
    area = 0.0
    n = len(contour)
    j = n - 1
    for i in range(n):
        area += (float(contour[j][0]) + contour[i][0]) * (float(contour[j][1]) - contour[i][1])
        j = i
    area = -area * 0.5
    if area < 0:
        contour = contour[::-1]
    if abs(delta) < 1e-20:
        return np.array(contour, dtype=np.int64)

    y = arc_tolerance if arc_tolerance > 0 else 0.25
    if y > abs(delta) * 0.25:
        y = abs(delta) * 0.25
    steps = math.pi / math.acos(1 - y / abs(delta))
    if steps > abs(delta) * math.pi:
        steps = abs(delta) * math.pi
    m_sin = math.sin(2 * math.pi / steps)
    m_cos = math.cos(2 * math.pi / steps)
    steps_per_rad = steps / (2 * math.pi)
    if delta < 0:
        m_sin = -m_sin

    def unit_normal(p1, p2):
        if p1 == p2:
            return (0.0, 0.0)
        dx = float(p2[0] - p1[0])
        dy = float(p2[1] - p1[1])
        f = 1.0 / math.sqrt(dx * dx + dy * dy)
        return (dy * f, -dx * f)

    normals = [unit_normal(contour[i], contour[(i + 1) % n]) for i in range(n)]
    dest: List[Tuple[int, int]] = []
    k = n - 1
    for j in range(n):
        sx, sy = contour[j]
        nk, nj = normals[k], normals[j]
        sin_a = nk[0] * nj[1] - nj[0] * nk[1]
        done = False
        if abs(sin_a * delta) < 1.0:
            cos_a = nk[0] * nj[0] + nj[1] * nk[1]
            if cos_a > 0:
                dest.append((_clipper_round(sx + nk[0] * delta), _clipper_round(sy + nk[1] * delta)))
                done = True
        elif sin_a > 1.0:
            sin_a = 1.0
        elif sin_a < -1.0:
            sin_a = -1.0
        if not done:
            if sin_a * delta < 0:
                dest.append((_clipper_round(sx + nk[0] * delta), _clipper_round(sy + nk[1] * delta)))
                dest.append((sx, sy))
                dest.append((_clipper_round(sx + nj[0] * delta), _clipper_round(sy + nj[1] * delta)))
            else:  # DoRound
                a = math.atan2(sin_a, nk[0] * nj[0] + nk[1] * nj[1])
                nsteps = max(_clipper_round(steps_per_rad * abs(a)), 1)
                X, Y = nk
                for _ in range(nsteps):
                    dest.append((_clipper_round(sx + X * delta), _clipper_round(sy + Y * delta)))
                    X2 = X
                    X = X * m_cos - m_sin * Y
                    Y = X2 * m_sin + Y * m_cos
                dest.append((_clipper_round(sx + nj[0] * delta), _clipper_round(sy + nj[1] * delta)))
        k = j
    return np.array(dest, dtype=np.int64)


# --------------------------------------------------------------------------- #
# DB post-process (DBPostProcess, score_mode='fast', box_type='quad', no dilation)
# --------------------------------------------------------------------------- #

def get_mini_boxes(contour: np.ndarray):
    bounding_box = cv2.minAreaRect(contour)
    points = sorted(list(cv2.boxPoints(bounding_box)), key=lambda p: p[0])
    if points[1][1] > points[0][1]:
        i1, i4 = 0, 1
    else:
        i1, i4 = 1, 0
    if points[3][1] > points[2][1]:
        i2, i3 = 2, 3
    else:
        i2, i3 = 3, 2
    box = [points[i1], points[i2], points[i3], points[i4]]
    return box, min(bounding_box[1])


def box_score_fast(bitmap: np.ndarray, _box: np.ndarray) -> float:
    h, w = bitmap.shape[:2]
    box = _box.copy()
    xmin = np.clip(np.floor(box[:, 0].min()).astype("int32"), 0, w - 1)
    xmax = np.clip(np.ceil(box[:, 0].max()).astype("int32"), 0, w - 1)
    ymin = np.clip(np.floor(box[:, 1].min()).astype("int32"), 0, h - 1)
    ymax = np.clip(np.ceil(box[:, 1].max()).astype("int32"), 0, h - 1)
    mask = np.zeros((ymax - ymin + 1, xmax - xmin + 1), dtype=np.uint8)
    box[:, 0] = box[:, 0] - xmin
    box[:, 1] = box[:, 1] - ymin
    cv2.fillPoly(mask, box.reshape(1, -1, 2).astype("int32"), 1)
    return cv2.mean(bitmap[ymin:ymax + 1, xmin:xmax + 1], mask)[0]


def polygon_area_length(box: np.ndarray) -> Tuple[float, float]:
    """shapely Polygon(box).area / .length for a simple ring (float64 shoelace / perimeter)."""
    p = np.asarray(box, dtype=np.float64)
    q = np.roll(p, -1, axis=0)
    area = abs(float(np.sum(p[:, 0] * q[:, 1] - q[:, 0] * p[:, 1])) * 0.5)
    length = float(np.sum(np.sqrt(((q - p) ** 2).sum(axis=1))))
    return area, length


def unclip(box: np.ndarray, unclip_ratio: float = 1.5) -> np.ndarray:
    area, length = polygon_area_length(box)
    distance = area * unclip_ratio / length
    return clipper_offset_round(box, distance)


def db_postprocess(pred: np.ndarray, shape, thresh: float = 0.3, box_thresh: float = 0.6,
                   max_candidates: int = 1000, unclip_ratio: float = 1.5, min_size: int = 3):
    """pred: float32 [rh, rw] probability map of one image. Returns (int32 [N,4,2], scores)."""
    src_h, src_w = int(shape[0]), int(shape[1])
    bitmap = pred > thresh
    height, width = bitmap.shape
    outs = cv2.findContours((bitmap * 255).astype(np.uint8), cv2.RETR_LIST, cv2.CHAIN_APPROX_SIMPLE)
    contours = outs[0] if len(outs) == 2 else outs[1]
    num_contours = min(len(contours), max_candidates)
    boxes, scores = [], []
    for index in range(num_contours):
        contour = contours[index]
        points, sside = get_mini_boxes(contour)
        if sside < min_size:
            continue
        points = np.array(points)
        score = box_score_fast(pred, points.reshape(-1, 2))
        if box_thresh > score:
            continue
        expanded = unclip(points, unclip_ratio)
        if len(expanded) < 3:
            continue
        box = expanded.reshape(-1, 1, 2).astype(np.int32)
        box, sside = get_mini_boxes(box)
        if sside < min_size + 2:
            continue
        box = np.array(box)
        box[:, 0] = np.clip(np.round(box[:, 0] / width * src_w), 0, src_w)
        box[:, 1] = np.clip(np.round(box[:, 1] / height * src_h), 0, src_h)
        boxes.append(box.astype("int32"))
        scores.append(score)
    return np.array(boxes, dtype="int32").reshape(-1, 4, 2), scores


def order_points_clockwise(pts: np.ndarray) -> np.ndarray:
    rect = np.zeros((4, 2), dtype="float32")
    s = pts.sum(axis=1)
    rect[0] = pts[np.argmin(s)]
    rect[2] = pts[np.argmax(s)]
    tmp = np.delete(pts, (np.argmin(s), np.argmax(s)), axis=0)
    diff = np.diff(np.array(tmp), axis=1)
    rect[1] = tmp[np.argmin(diff)]
    rect[3] = tmp[np.argmax(diff)]
    return rect


def filter_tag_det_res(dt_boxes: np.ndarray, image_shape) -> np.ndarray:
    img_height, img_width = image_shape[0:2]
    out = []
    for box in dt_boxes:
        box = np.array(box)
        box = order_points_clockwise(box)
        for p in range(box.shape[0]):
            box[p, 0] = int(min(max(box[p, 0], 0), img_width - 1))
            box[p, 1] = int(min(max(box[p, 1], 0), img_height - 1))
        rect_width = int(np.linalg.norm(box[0] - box[1]))
        rect_height = int(np.linalg.norm(box[0] - box[3]))
        if rect_width <= 3 or rect_height <= 3:
            continue
        out.append(box)
    return np.array(out, dtype=np.float32).reshape(-1, 4, 2) if out else np.zeros((0, 4, 2), dtype=np.float32)


# --------------------------------------------------------------------------- #
# TextSystem glue: box ordering, crops
# --------------------------------------------------------------------------- #

def sorted_boxes(dt_boxes: np.ndarray) -> List[np.ndarray]:
    num_boxes = dt_boxes.shape[0]
    sb = sorted(dt_boxes, key=lambda x: (x[0][1], x[0][0]))
    _boxes = list(sb)
    for i in range(num_boxes - 1):
        for j in range(i, -1, -1):
            if abs(_boxes[j + 1][0][1] - _boxes[j][0][1]) < 10 and (_boxes[j + 1][0][0] < _boxes[j][0][0]):
                _boxes[j], _boxes[j + 1] = _boxes[j + 1], _boxes[j]
            else:
                break
    return _boxes


def get_rotate_crop_image(img: np.ndarray, points: np.ndarray) -> np.ndarray:
    assert len(points) == 4
    cw = int(max(np.linalg.norm(points[0] - points[1]), np.linalg.norm(points[2] - points[3])))
    ch = int(max(np.linalg.norm(points[0] - points[3]), np.linalg.norm(points[1] - points[2])))
    pts_std = np.float32([[0, 0], [cw, 0], [cw, ch], [0, ch]])
    M = cv2.getPerspectiveTransform(points, pts_std)
    dst = cv2.warpPerspective(img, M, (cw, ch), borderMode=cv2.BORDER_REPLICATE, flags=cv2.INTER_CUBIC)
    dh, dw = dst.shape[0:2]
    if dh * 1.0 / dw >= 1.5:
        dst = np.rot90(dst)
    return dst


# --------------------------------------------------------------------------- #
# recognition pre-process (TextRecognizer.__call__ batching + resize_norm_img)
# --------------------------------------------------------------------------- #

def rec_batches(wh_ratios: Sequence[float], rec_batch_num: int = 6, img_h: int = 48, img_w: int = 320):
    """Yield (indices_in_original_order, imgW) per upstream batch (aspect-sorted, <= rec_batch_num)."""
    order = np.argsort(np.array(wh_ratios))
    n = len(wh_ratios)
    for beg in range(0, n, rec_batch_num):
        end = min(n, beg + rec_batch_num)
        max_wh_ratio = img_w / img_h
        for ino in range(beg, end):
            max_wh_ratio = max(max_wh_ratio, wh_ratios[order[ino]])
        yield [int(order[i]) for i in range(beg, end)], int(img_h * max_wh_ratio)


def resize_norm_img(img: np.ndarray, img_w: int, img_h: int = 48) -> np.ndarray:
    """crop (BGR u8) -> float32 [3, img_h, img_w], resized to H, right zero-padded."""
    h, w = img.shape[:2]
    ratio = w / float(h)
    if math.ceil(img_h * ratio) > img_w:
        resized_w = img_w
    else:
        resized_w = int(math.ceil(img_h * ratio))
    resized = cv2.resize(img, (resized_w, img_h))
    x = resized.astype("float32").transpose((2, 0, 1)) / 255
    x -= 0.5
    x /= 0.5
    out = np.zeros((3, img_h, img_w), dtype=np.float32)
    out[:, :, :resized_w] = x
    return out


# --------------------------------------------------------------------------- #
# CTC greedy decode (CTCLabelDecode)
# --------------------------------------------------------------------------- #

def ctc_decode_ids(probs: np.ndarray) -> Tuple[List[int], float, List[float]]:
    """probs [T, C] -> (kept class ids, mean kept max-prob (0 if none), kept probs)."""
    idx = probs.argmax(axis=1)
    prob = probs.max(axis=1)
    sel = np.ones(len(idx), dtype=bool)
    sel[1:] = idx[1:] != idx[:-1]
    sel &= idx != 0
    kept = prob[sel]
    if len(kept) == 0:
        return [], 0.0, []
    return [int(i) for i in idx[sel]], float(np.mean(kept)), [float(p) for p in kept]


def ids_to_text(ids: Sequence[int], characters: Sequence[str]) -> str:
    return "".join(characters[i] for i in ids)
