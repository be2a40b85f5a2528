"""ctypes binding of libvse_b200.so (include/vse_b200.h).

This is the thin host side the north star asks for: Python owns no arithmetic, only buffers.
There is NO CPU fallback: if the library or a CUDA device is missing, construction raises.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np

_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libvse_b200.so")
_lib = None

PLAN_DET, PLAN_REC = 0, 1
MEM_HOST, MEM_PINNED, MEM_DEVICE = 0, 1, 2
PRECISION_FP16, PRECISION_FP32, PRECISION_TF32, PRECISION_FP32_TC = 0, 1, 2, 3
FLAG_NO_TENSOR_CORES = 1
FLAG_NO_FAST_KERNELS = 2
FLAG_NO_FUSED_HEAD, FLAG_NO_SE_FUSION, FLAG_NO_ROWBOX, FLAG_NO_FAST_DW, FLAG_NO_FAST_STEM, FLAG_NO_PIXEL_PACK = 4, 8, 16, 32, 64, 128
FLAG_NO_CONCAT_GATHER, FLAG_NO_HALO = 256, 512
FLAG_DET_FP32, FLAG_DET_TF32 = 1024, 2048
FLAG_NO_SE_CONV = 4096
FLAG_DET_FP32_TC = 8192
FLAG_DWPW_FUSION = 16384


class VseConfig(C.Structure):
    _fields_ = [("device", C.c_int32), ("precision", C.c_int32), ("det_limit_side_len", C.c_int32),
                ("det_thresh", C.c_float), ("det_box_thresh", C.c_float), ("det_unclip_ratio", C.c_float),
                ("det_max_candidates", C.c_int32), ("rec_image_h", C.c_int32), ("rec_image_w", C.c_int32),
                ("rec_batch_num", C.c_int32), ("max_boxes_per_frame", C.c_int32), ("flags", C.c_int32)]


class VseResult(C.Structure):
    _fields_ = [("box_capacity", C.c_int32), ("max_text_len", C.c_int32), ("n_boxes", C.POINTER(C.c_int32)),
                ("quads", C.POINTER(C.c_float)), ("det_score", C.POINTER(C.c_float)), ("ids", C.POINTER(C.c_int32)),
                ("id_len", C.POINTER(C.c_int32)), ("rec_score", C.POINTER(C.c_float)),
                ("rec_width", C.POINTER(C.c_int32)), ("timings_ms", C.c_float * 8)]


EXPORTS = ["vse_default_config", "vse_abi_version", "vse_device_count", "vse_create", "vse_destroy", "vse_last_error",
           "vse_load_plan", "vse_set_conv_input_ranges", "vse_run", "vse_det_only", "vse_prefetch", "vse_launch_count", "vse_tc_launch_count", "vse_debug_run_plan", "vse_debug_get_value",
           "vse_debug_resize_bilinear", "vse_debug_db_postprocess", "vse_debug_crop", "vse_debug_time_steps"]


def load_library(path: Optional[str] = None):
    """dlopen the engine. Raises if it has not been built (python __graft_entry__.py)."""
    global _lib
    if _lib is not None:
        return _lib
    path = path or _LIB_PATH
    if not os.path.exists(path):
        raise RuntimeError(f"{path} is missing: build it with `python __graft_entry__.py` (there is no CPU fallback)")
    lib = C.CDLL(path)
    vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
    pp_u8 = C.POINTER(C.c_void_p)
    p_i32 = C.POINTER(C.c_int32)
    p_f32 = C.POINTER(C.c_float)
    lib.vse_default_config.argtypes = [C.POINTER(VseConfig)]
    lib.vse_default_config.restype = None
    lib.vse_abi_version.restype = C.c_int
    lib.vse_device_count.restype = C.c_int
    lib.vse_create.argtypes = [C.POINTER(VseConfig), C.POINTER(vp)]
    lib.vse_destroy.argtypes = [vp]
    lib.vse_destroy.restype = None
    lib.vse_last_error.argtypes = [vp]
    lib.vse_last_error.restype = C.c_char_p
    lib.vse_load_plan.argtypes = [vp, i32, vp, C.c_size_t]
    lib.vse_set_conv_input_ranges.argtypes = [vp, i32, p_f32, i32]
    for fn in (lib.vse_run, lib.vse_det_only):
        fn.argtypes = [vp, pp_u8, p_i32, p_i32, p_i32, i32, i32, C.POINTER(VseResult)]
    lib.vse_prefetch.argtypes = [vp, pp_u8, p_i32, p_i32, p_i32, i32, i32]
    lib.vse_launch_count.argtypes = [vp]
    lib.vse_launch_count.restype = i64
    lib.vse_tc_launch_count.argtypes = [vp]
    lib.vse_tc_launch_count.restype = i64
    lib.vse_debug_run_plan.argtypes = [vp, i32, pp_u8, i32, i32, p_i32, p_i32, i32]
    lib.vse_debug_get_value.argtypes = [vp, i32, i32, p_f32, i64, p_i32]
    lib.vse_debug_get_value.restype = i64
    lib.vse_debug_resize_bilinear.argtypes = [vp, vp, i32, i32, i32, vp, i32, i32]
    lib.vse_debug_db_postprocess.argtypes = [vp, p_f32, i32, i32, i32, i32, p_f32, p_f32, i32, p_i32]
    lib.vse_debug_crop.argtypes = [vp, vp, i32, i32, p_f32, vp, i32, p_i32, p_i32]
    lib.vse_debug_time_steps.argtypes = [vp, i32, i32, p_f32, C.POINTER(C.c_int64), i32]
    _lib = lib
    return lib


def bench_mode() -> dict:
    """Engine keyword arguments of the mode bench.py times (BASELINE configs[1]); tests/test_gpu_real_video.py holds exactly
    this mode to the north-star bars (IoU >= 0.99 box-for-box, CER <= 1e-3) on the reference's sample videos.  fp16
    activation storage is ~1.5x faster but moves the 0.3 threshold crossing of the detector on ~1 % of real-video boxes
    (tests/test_gpu_real_video.py reports it), so it is not the mode that is timed."""
    return dict(precision=PRECISION_FP32_TC)


def mixed_mode() -> dict:
    """Detector in the fp32 tensor-core mode (its boxes hang on the 0.3 threshold crossing of the probability map), recogniser
    with fp16 activations (its bar is CER <= 1e-3 on the class ids; tests/test_gpu_real_video.py measures it on the reference's
    sample videos as mode "mixed")."""
    return dict(precision=PRECISION_FP16, flags=FLAG_DET_FP32_TC)


def accurate_mode() -> dict:
    """Engine keyword arguments for the accurate-mode models (V4/ch_det + V4/<lang>_rec, reference
    backend/tools/paddle_model_config.py:69-71).  V4/ch_det's activations exceed the fp16 range (LK-PAN outputs reach 1.7e5);
    the fp32 tensor-core mode scales every convolution's operands by its calibrated power of two (calibration/V4__ch_det.json),
    so the same mode serves it."""
    return dict(precision=PRECISION_FP32_TC)


def device_count() -> int:
    return int(load_library().vse_device_count())


@dataclass
class FrameResult:
    quads: np.ndarray        # float32 [n,4,2]
    det_scores: np.ndarray   # float32 [n]
    ids: List[List[int]]     # CTC class ids per box (empty for det-only)
    rec_scores: np.ndarray   # float32 [n]
    rec_widths: np.ndarray   # int32 [n]


def _i32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.int32)


class Engine:
    """One engine per (process, GPU) — mirrors the lifetime of PaddleOCR(...) in reference ocr.py:88-113."""

    def __init__(self, device: int = 0, precision: int = PRECISION_FP32_TC, rec_image_h: int = 48, rec_image_w: int = 320,
                 rec_batch_num: int = 6, det_limit_side_len: int = 960, det_thresh: float = 0.3,
                 det_box_thresh: float = 0.6, det_unclip_ratio: float = 1.5, max_boxes_per_frame: int = 64,
                 max_text_len: int = 256, flags: int = 0):
        self.lib = load_library()
        cfg = VseConfig()
        self.lib.vse_default_config(C.byref(cfg))
        cfg.device, cfg.precision = device, precision
        cfg.rec_image_h, cfg.rec_image_w, cfg.rec_batch_num = rec_image_h, rec_image_w, rec_batch_num
        cfg.det_limit_side_len, cfg.det_thresh = det_limit_side_len, det_thresh
        cfg.det_box_thresh, cfg.det_unclip_ratio = det_box_thresh, det_unclip_ratio
        cfg.max_boxes_per_frame = max_boxes_per_frame
        cfg.flags = flags
        self.cfg = cfg
        self.max_text_len = max_text_len
        self._rows_per_frame = max(1, max_boxes_per_frame)     # result rows offered per frame (grown on VSE_ERR_CAPACITY)
        self._h = C.c_void_p()
        rc = self.lib.vse_create(C.byref(cfg), C.byref(self._h))
        if rc != 0:
            raise RuntimeError(f"vse_create failed ({rc}): {self.lib.vse_last_error(None).decode()}")
        self.plan_names = {}

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self.lib.vse_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int, what: str):
        if rc < 0:
            raise RuntimeError(f"{what} failed ({rc}): {self.lib.vse_last_error(self._h).decode()}")

    def load_plan(self, which: int, blob: bytes, name: str = ""):
        buf = (C.c_char * len(blob)).from_buffer_copy(blob)
        self._check(self.lib.vse_load_plan(self._h, which, C.cast(buf, C.c_void_p), len(blob)), "vse_load_plan")
        self.plan_names[which] = name
        self._last_blob = blob
        if name:
            self._apply_calibration(which, name)

    def _apply_calibration(self, which: int, name: str):
        """Calibrated per-step input ranges (calibration/<model>.json, tools/calibrate_ranges.py) for the fp32 tensor-core mode."""
        import json
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "calibration",
                            "__".join(name.replace("\\", "/").rstrip("/").split("/")[-2:]) + ".json")
        if not os.path.exists(path):
            return
        with open(path) as f:
            cal = json.load(f)
        from . import plan as _P
        n_steps = _P.blob_step_count(self._last_blob)
        if int(cal["n_steps"]) != n_steps:
            import warnings
            warnings.warn(f"vse_b200: calibration file {os.path.basename(path)} was made for a plan with {cal['n_steps']} steps, this "
                          f"plan has {n_steps}: re-run tools/calibrate_ranges.py (default operand scales are used)", RuntimeWarning)
            return
        arr = np.zeros(int(cal["n_steps"]), np.float32)
        for k, v in cal["conv_input_absmax"].items():
            arr[int(k)] = v
        self._check(self.lib.vse_set_conv_input_ranges(self._h, which, arr.ctypes.data_as(C.POINTER(C.c_float)), len(arr)),
                    "vse_set_conv_input_ranges")

    @property
    def tc_launch_count(self) -> int:
        return int(self.lib.vse_tc_launch_count(self._h))

    @property
    def launch_count(self) -> int:
        return int(self.lib.vse_launch_count(self._h))

    # -- hot path --------------------------------------------------------- #
    def _run(self, frames: Sequence, heights, widths, strides, mem_kind: int, det_only: bool) -> Tuple[List[FrameResult], np.ndarray]:
        """One vse_run / vse_det_only call.  The result buffers are caller-owned (C-ABI contract); when the engine reports
        VSE_ERR_CAPACITY (more boxes than rows, or a line longer than max_text_len symbols) they are grown and the call is
        repeated — upstream keeps up to 1000 candidates per frame (DBPostProcess max_candidates) and never fails on a dense
        frame, so neither does this binding."""
        n = len(frames)
        ptrs = (C.c_void_p * max(n, 1))(*[int(p) for p in frames])
        h, w = _i32(heights), _i32(widths)
        st = _i32(strides) if strides is not None else None
        as_p = lambda a, t: a.ctypes.data_as(C.POINTER(t))
        fn = self.lib.vse_det_only if det_only else self.lib.vse_run
        for attempt in range(8):
            cap = max(1, n * int(self._rows_per_frame))
            T = self.max_text_len
            # keep one set of buffers per engine and batch size instead of allocating and zero-filling ~2 MB per call — the
            # engine writes every field it reports a count for
            key = (n, cap, T)
            if getattr(self, "_res_key", None) != key:
                self._res_key = key
                self._res_buf = (np.zeros(max(n, 1), np.int32), np.zeros((cap, 4, 2), np.float32), np.zeros(cap, np.float32),
                                 np.zeros((cap, T), np.int32), np.zeros(cap, np.int32), np.zeros(cap, np.float32),
                                 np.zeros(cap, np.int32))
            n_boxes, quads, det_score, ids, id_len, rec_score, rec_width = self._res_buf
            n_boxes[:] = 0
            res = VseResult()
            res.box_capacity, res.max_text_len = cap, T
            res.n_boxes, res.quads, res.det_score = as_p(n_boxes, C.c_int32), as_p(quads, C.c_float), as_p(det_score, C.c_float)
            res.ids, res.id_len, res.rec_score = as_p(ids, C.c_int32), as_p(id_len, C.c_int32), as_p(rec_score, C.c_float)
            res.rec_width = as_p(rec_width, C.c_int32)
            rc = fn(self._h, ptrs, as_p(h, C.c_int32), as_p(w, C.c_int32), as_p(st, C.c_int32) if st is not None else None, n,
                    mem_kind, C.byref(res))
            if rc == -3 and attempt < 7:      # VSE_ERR_CAPACITY
                msg = self.lib.vse_last_error(self._h).decode()
                if "max_text_len" in msg and self.max_text_len < 8192:
                    self.max_text_len *= 4
                    continue
                if "box_capacity" in msg and self._rows_per_frame < 1000:
                    self._rows_per_frame = min(1000, self._rows_per_frame * 4)
                    continue
            break
        self._check(rc, "vse_det_only" if det_only else "vse_run")
        out: List[FrameResult] = []
        row = 0
        for f in range(n):
            k = int(n_boxes[f])
            sl = slice(row, row + k)
            out.append(FrameResult(quads[sl].copy(), det_score[sl].copy(),
                                   [ids[r, :id_len[r]].tolist() for r in range(row, row + k)],
                                   rec_score[sl].copy(), rec_width[sl].copy()))
            row += k
        return out, np.array(list(res.timings_ms), dtype=np.float32)

    def run(self, frames: Sequence[np.ndarray], det_only: bool = False) -> List[FrameResult]:
        """frames: BGR uint8 HWC numpy arrays (host memory)."""
        frames = [np.ascontiguousarray(f) if not f.flags["C_CONTIGUOUS"] else f for f in frames]
        for f in frames:
            if f.dtype != np.uint8 or f.ndim != 3 or f.shape[2] != 3:
                raise ValueError("frames must be uint8 HxWx3 (BGR)")
        res, self.last_timings = self._run([f.ctypes.data for f in frames], [f.shape[0] for f in frames],
                                           [f.shape[1] for f in frames], [f.strides[0] for f in frames], MEM_HOST, det_only)
        return res

    def prefetch(self, ptrs: Sequence[int], heights, widths, strides=None, mem_kind: int = MEM_PINNED):
        """Start the host->device copy of the NEXT batch; the following run_device(...) with the same pointers uses it."""
        n = len(ptrs)
        if n == 0:
            return
        arr = (C.c_void_p * n)(*[int(p) for p in ptrs])
        h, w = _i32(heights), _i32(widths)
        st = _i32(strides) if strides is not None else None
        as_p = lambda a: a.ctypes.data_as(C.POINTER(C.c_int32))
        self._check(self.lib.vse_prefetch(self._h, arr, as_p(h), as_p(w), as_p(st) if st is not None else None, n, mem_kind),
                    "vse_prefetch")

    def run_device(self, ptrs: Sequence[int], heights, widths, strides=None, det_only: bool = False, mem_kind: int = MEM_DEVICE):
        """frames already resident (device pointers, e.g. torch tensors' data_ptr()) or pinned host pointers."""
        res, self.last_timings = self._run(ptrs, heights, widths, strides, mem_kind, det_only)
        return res

    # -- test hooks -------------------------------------------------------- #
    def debug_run_plan(self, which: int, images_bgrx: Sequence[np.ndarray], valid_w: Optional[Sequence[int]] = None,
                       keep_all: bool = True):
        imgs = [np.ascontiguousarray(im, dtype=np.uint8) for im in images_bgrx]
        h = imgs[0].shape[0]
        assert all(im.shape[0] == h and im.shape[2] == 4 for im in imgs)
        w = _i32([im.shape[1] for im in imgs])
        vw = _i32(valid_w if valid_w is not None else [im.shape[1] for im in imgs])
        ptrs = (C.c_void_p * len(imgs))(*[im.ctypes.data for im in imgs])
        rc = self.lib.vse_debug_run_plan(self._h, which, ptrs, len(imgs), h, w.ctypes.data_as(C.POINTER(C.c_int32)),
                                         vw.ctypes.data_as(C.POINTER(C.c_int32)), int(keep_all))
        self._check(rc, "vse_debug_run_plan")

    def debug_get_value(self, which: int, vid: int) -> Optional[np.ndarray]:
        ch = C.c_int32(0)
        n = self.lib.vse_debug_get_value(self._h, which, vid, None, 0, C.byref(ch))
        self._check(n, "vse_debug_get_value")
        if n == 0:
            return None
        out = np.empty(n, np.float32)
        n2 = self.lib.vse_debug_get_value(self._h, which, vid, out.ctypes.data_as(C.POINTER(C.c_float)), n, C.byref(ch))
        self._check(n2, "vse_debug_get_value")
        return out.reshape(-1, ch.value)

    def debug_time_steps(self, which: int, reps: int = 5):
        """-> (ms per step, info[n,8]) for the last run of plan `which` (see include/vse_b200.h)."""
        cap = 1024
        ms = np.zeros(cap, np.float32)
        info = np.zeros((cap, 8), np.int64)
        n = self.lib.vse_debug_time_steps(self._h, which, reps, ms.ctypes.data_as(C.POINTER(C.c_float)),
                                          info.ctypes.data_as(C.POINTER(C.c_int64)), cap)
        self._check(n, "vse_debug_time_steps")
        return ms[:n].copy(), info[:n].copy()

    def debug_resize(self, img_bgr: np.ndarray, dh: int, dw: int) -> np.ndarray:
        img = np.ascontiguousarray(img_bgr, dtype=np.uint8)
        out = np.empty((dh, dw, 4), np.uint8)
        rc = self.lib.vse_debug_resize_bilinear(self._h, img.ctypes.data, img.shape[0], img.shape[1], img.strides[0],
                                                out.ctypes.data, dh, dw)
        self._check(rc, "vse_debug_resize_bilinear")
        return out

    def debug_db_postprocess(self, prob: np.ndarray, src_h: int, src_w: int, capacity: int = 256):
        prob = np.ascontiguousarray(prob, dtype=np.float32)
        quads = np.zeros((capacity, 4, 2), np.float32)
        scores = np.zeros(capacity, np.float32)
        n = C.c_int32(0)
        rc = self.lib.vse_debug_db_postprocess(self._h, prob.ctypes.data_as(C.POINTER(C.c_float)), prob.shape[0], prob.shape[1],
                                               src_h, src_w, quads.ctypes.data_as(C.POINTER(C.c_float)),
                                               scores.ctypes.data_as(C.POINTER(C.c_float)), capacity, C.byref(n))
        self._check(rc, "vse_debug_db_postprocess")
        return quads[:n.value].copy(), scores[:n.value].copy()

    def debug_crop(self, frame: np.ndarray, quad: np.ndarray) -> np.ndarray:
        frame = np.ascontiguousarray(frame, dtype=np.uint8)
        quad = np.ascontiguousarray(quad, dtype=np.float32)
        cap = 4 * 1024 * 1024 * 3
        out = np.empty(cap, np.uint8)
        oh, ow = C.c_int32(0), C.c_int32(0)
        rc = self.lib.vse_debug_crop(self._h, frame.ctypes.data, frame.shape[0], frame.shape[1],
                                     quad.ctypes.data_as(C.POINTER(C.c_float)), out.ctypes.data, cap, C.byref(oh), C.byref(ow))
        self._check(rc, "vse_debug_crop")
        return out[:oh.value * ow.value * 3].reshape(oh.value, ow.value, 3).copy()
