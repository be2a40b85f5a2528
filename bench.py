#!/usr/bin/env python
"""bench.py — OCR frames/s (det + rec) of the B200 engine on BASELINE.json configs[1]:
synthetic 1080p subtitle frames, batch = 32 frames per vse_run call, V4/ch_det_fast + V4/en_rec_fast.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

A "step" is one pass of the hot path (vse_run: det pre-process, det net, DB post-process, crops, rec net, CTC decode)
over one batch of 32 frames per GPU.  `value` times the steps with the frames already resident in HBM; `e2e` times the
same call with the frames in pinned HOST memory (H2D of the frames and D2H of the results inside the timed region).
Weak scaling: every rank owns its own contiguous 32-frame range of each global batch (no collective on the hot path;
one NCCL broadcast of the packed plans at start-up).  Prints ONE JSON line on rank 0.

`--impl reference` times the CPU restatement of the reference path (oracle/: torch-CPU fp32 graph arithmetic + cv2
host logic, one frame per call, det batch 1, rec batches of <= 6 — exactly how the reference drives paddleocr) on the
host cores.  The reference's own Paddle runtime is a pip dependency that is absent from this image.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DET, REC = "V4/ch_det_fast", "V4/en_rec_fast"
METRIC = "ocr_frames_per_sec_det_rec"
UNIT = "frames/s"
FALLBACK_HBM_GBS = 6650.0   # /opt/skills/guides/B200_PROFILING.md fallback when MEASURED_PEAKS.json is absent


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=32, help="frames per vse_run call per GPU")
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--pool", type=int, default=3, help="distinct batches cycled through (3 x 32 x 6.2 MB > L2)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="CPU seconds spent on the cpu_baseline leg (>= one pass over a step's frames)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--det", default=None, help="detector model (default V4/ch_det_fast = BASELINE configs[1]); needs its packed plan")
    ap.add_argument("--rec", default=None, help="recogniser model (default V4/en_rec_fast)")
    ap.add_argument("--flags", type=int, default=0, help="vse_config.flags (VSE_FLAG_* A/B switches, profiling only)")
    ap.add_argument("--precision", default=None, choices=["fp16", "fp32", "tf32", "fp32_tc", "mixed"],
                    help="activation / product precision (default: engine.bench_mode(), the mode held to the parity bar)")
    ap.add_argument("--frame-stride", type=int, default=7, help="synthetic stream index step between consecutive frames of the pool")
    ap.add_argument("--config", type=int, default=1, choices=[1, 2, 3, 4],
                    help="BASELINE.json configs[k]: 1 = the headline (default); 2 = test_cn.mp4 accurate mode, V4/ch_det + V4/ch_rec, "
                         "batch 64; 3 = four sample videos, fast mode, frame ranges sharded over the ranks (whole jobs: decode -> "
                         ".srt); 4 = 4K synthetic 50k-frame stream, accurate models, frame-parallel")
    ap.add_argument("--video", default=None, help="take the frames from this video (every --frame-stride-th frame) instead of the synthetic stream")
    args = ap.parse_args()
    vids = os.path.join(ROOT, "tests", "golden", "_videos")
    if args.config == 2:
        args.det, args.rec = args.det or "V4/ch_det", args.rec or "V4/ch_rec"
        args.batch, args.height, args.width, args.pool = 64, 1080, 1440, 2
        if args.video is None and os.path.exists(os.path.join(vids, "test_cn.mp4")):
            args.video = os.path.join(vids, "test_cn.mp4")
        args.frame_stride = 3
    elif args.config == 3:
        if "--steps" not in sys.argv:
            args.steps = 1
        if "--warmup" not in sys.argv:
            args.warmup = 0
    elif args.config == 4:
        args.det, args.rec = args.det or "V4/ch_det", args.rec or "V4/ch_rec"
        args.batch, args.height, args.width, args.pool = 16, 2160, 3840, 2
        if "--steps" not in sys.argv:
            args.steps = -(-50000 // (max(args.gpus, 1) * args.batch))      # the whole 50k-frame stream, sharded
    return args


def apply_model_args(args):
    global DET, REC
    if args.det:
        DET = args.det
    if args.rec:
        REC = args.rec


def frame_index(p: int, k: int, world_b: int, stride: int, per_rank: int = 0) -> int:
    """Index into the synthetic stream of frame k of global batch p.  Consecutive frames of a rank are `stride` stream frames
    apart, so a pool of 3 x 32 frames walks over ~11 subtitles (45 frames of text + 15 blank each), not 2.  Weak scaling wants
    the SAME work on every GPU at every N: rank r (= k // per_rank) gets the N = 1 frame set shifted by r stream frames — the same
    subtitles, other frames (noise, background phase).  Cutting ONE stream into contiguous ranges (what shard.frame_range does for
    a real job) hands the ranks different text densities (26-41 lines per step at N = 8, profiles/r02_bench_n8.json) and the
    max-over-ranks time then measures the stream's imbalance, not the system."""
    per_rank = per_rank or world_b
    r, slot = divmod(k, per_rank)
    return (p * per_rank + slot) * stride + r


def workload_config(args, world: int):
    """The `config` object both arms print (the driver compares them)."""
    B, H, W = args.batch, args.height, args.width
    src = f"frames of {os.path.basename(args.video)} (reference sample video)" if args.video else f"synthetic {H}p subtitle frames (SURVEY.md §8d generator)"
    return {"workload": f"{src}, {DET} + {REC}", "baseline_config": args.config,
            "frames_per_step_per_gpu": B, "global_frames_per_step": world * B, "frame": [H, W, 3],
            "parallelism": f"frame-range sharding x{world}",
            "frames": f"rank r, slot k of pool batch p: stream index (p * {B} + k) * {args.frame_stride} + r (every rank gets the N = 1 frame "
                      f"set shifted by r stream frames), {args.pool} pool batches cycled",
            "l2": f"inputs larger than L2: {args.pool} distinct batches x {B * H * W * 3 / 1e6:.0f} MB cycled"}


def video_frames_at(path: str, indices):
    """{index: frame} for the given 0-based frame indices of a video, one sequential decode."""
    import cv2
    cap = cv2.VideoCapture(path)
    need, out, no = set(indices), {}, 0
    while need and no <= max(need):
        ok, fr = cap.read()
        if not ok:
            break
        if no in need:
            out[no] = fr
        no += 1
    cap.release()
    return out


def env_rank():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


# --------------------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 6:
                continue
            try:
                sm.append(float(p[0])); mx.append(float(p[1]))
            except ValueError:
                continue
            for n, v in zip(names, p[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------------------
# CPU arm (oracle) — used by `--impl reference` and by the cpu_baseline leg
# --------------------------------------------------------------------------------------------------
def cpu_oracle(det_blob, rec_blob):
    import torch
    from oracle.pipeline import OraclePipeline      # checker / CPU baseline only
    torch.set_num_threads(os.cpu_count() or 1)
    return OraclePipeline.from_plans(det_blob, rec_blob), torch.get_num_threads()


def time_cpu(oracle, frames, budget_s):
    """frames/s of the CPU restatement: cycles through `frames` (one step's batch) until the budget of CPU seconds is
    spent — at least one full pass; the first call is an untimed warm-up."""
    oracle.ocr(frames[0])
    t0 = time.perf_counter()
    n = 0
    while True:
        for f in frames:
            oracle.ocr(f)
            n += 1
        if time.perf_counter() - t0 >= budget_s:
            break
    dt = time.perf_counter() - t0
    return n / dt, n, dt


def run_reference(args, rank, world):
    if rank != 0:
        return
    from video_subtitle_extractor_b200 import weights
    from video_subtitle_extractor_b200.synth import SynthStream
    det_blob, rec_blob = weights.load_plan_blob(DET), weights.load_plan_blob(REC)
    oracle, cores = cpu_oracle(det_blob, rec_blob)
    per_step = 4 if args.config == 1 else 1          # bounded sample of the batch per step (server graphs: ~12 s per frame)
    stream = SynthStream(args.height, args.width)
    # the SAME frames the B200 arm times: frames 0, 8, 16, 24 of pool batch (step mod pool) of rank 0
    B = args.batch
    pick = [k * (B // per_step) for k in range(per_step)]
    if args.video:
        vf = video_frames_at(args.video, [frame_index(p, k, world * B, args.frame_stride, B) for p in range(args.pool) for k in pick])
        frames = {(p, k): vf[frame_index(p, k, world * B, args.frame_stride, B)] for p in range(args.pool) for k in pick}
        args.height, args.width = next(iter(vf.values())).shape[:2]
    else:
        frames = {(p, k): stream.frame(frame_index(p, k, world * B, args.frame_stride, B)) for p in range(args.pool) for k in pick}
    for _ in range(max(args.warmup, 1)):
        oracle.ocr(frames[(0, 0)])
    t0 = time.perf_counter()
    for s in range(args.steps):
        for k in pick:
            oracle.ocr(frames[(s % args.pool, k)])
    dt = time.perf_counter() - t0
    fps = args.steps * per_step / dt
    sample = (f"frames {pick} of the {B} frames of each step's batch ({args.height}x{args.width}), one frame per call, rec batches <= 6")
    emit(({
        "impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, world),
        "arm": "CPU restatement of the reference path (torch-CPU fp32 + cv2, all host threads), bounded sample: "
               f"{per_step} of the {args.batch} frames of each step",
        "cpu_baseline": {"value": fps, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# --------------------------------------------------------------------------------------------------
# B200 arm
# --------------------------------------------------------------------------------------------------
def step_roofline(eng, E):
    """Per-kernel achieved HBM GB/s of the dominant kernel, from CUDA events between plan steps (engine stream)."""
    from video_subtitle_extractor_b200 import plan as P
    rows, rows_out, rows_op = [], [], []
    for which in (E.PLAN_DET, E.PLAN_REC):
        try:
            ms, info = eng.debug_time_steps(which, reps=5)
        except RuntimeError:
            continue
        for t, i in zip(ms, info):
            opk, pin, pout, cin, cout, taps, eb_in, eb_out = (int(v) for v in i)
            op, kind = opk & 0xFF, opk >> 8
            if kind == 3:          # ran inside the previous step's fused kernel: fold its output bytes into that row
                w_, n_, t_, b_, f_ = rows[-1]
                if n_ == "conv_tc_kernel" and rows_op[-1] != P.OP_DWCONV:
                    # a 1x1 conv heading a fused squeeze-excite group: the step's time covers three launches (pool of the conv
                    # input, gate, conv with the gate in its epilogue) — kept apart from the plain conv launches
                    n_ = "se_conv_group(conv_tc+gpool+gate)"
                prev_out = rows_out.pop()
                rows[-1] = (w_, n_, t_ + float(t), b_ - prev_out + pout * eb_out + taps * cin * cout * 4, f_ + 2 * pout * cin * cout)
                rows_out.append(pout * eb_out)
                continue
            pad8 = lambda c: (c + 7) // 8 * 8
            wbytes = taps * cin * cout * 4 if op in (P.OP_CONV, P.OP_STEM, P.OP_DECONV2) else taps * cin * 4
            if kind == 1 and eb_in == 2:
                wbytes //= 2       # fp16 weight matrix (fp32 activations: tf32 words or fp16 hi + lo, 4 bytes per weight)
            out_bytes = pout * (cout if eb_out == 4 else pad8(cout)) * eb_out
            if op == P.OP_STEM:
                bytes_ = pin * 4 + pout * pad8(cout) * eb_out + wbytes
            elif op in (P.OP_CONV, P.OP_DWCONV, P.OP_DECONV2):
                bytes_ = pin * pad8(cin) * eb_in + out_bytes + wbytes
            else:
                bytes_ = pin * pad8(cin) * eb_in + pout * pad8(cout) * eb_out
            flops = 2 * pout * cin * cout * taps if op in (P.OP_CONV, P.OP_STEM) else (2 * pout * cin * cout if op == P.OP_DECONV2 else 0)
            name = P.OP_NAMES[op]
            if op == P.OP_CONV:
                kname = "conv_tc_kernel" if kind == 1 else "conv_simt_kernel"
            elif op == P.OP_STEM:
                kname = "stem_fast_kernel" if kind == 2 else "conv_simt_kernel"
            elif op == P.OP_DWCONV:
                # the engine's default depthwise family is the register-tiled kernel (engine.cu, VSE_DW_MODE switches it)
                dw_fast = {"0": "dwconv_fast_kernel", "1": "dwconv_tile_kernel"}.get(os.environ.get("VSE_DW_MODE", "2"), "dwconv_reg_kernel")
                # kind 1: computed inside the following 1x1 convolution's tensor-core kernel (fused depthwise -> pointwise); the 1x1
                # step follows as kind 3 and folds its output bytes, weights and FLOPs into this row
                kname = "conv_tc_kernel" if kind == 1 else dw_fast if kind == 2 else "dwconv_kernel"
            elif op == P.OP_DECONV2:
                kname = "db_head_fused_kernel" if kind == 2 else ("conv_tc_kernel" if kind == 1 else "deconv2_kernel")
            elif op == P.OP_LSTM:
                kname = "lstm_recurrent_kernel"
            else:
                kname = name.lower() + "_kernel"
            rows.append((which, kname, float(t), bytes_, flops))
            rows_op.append(op)
            rows_out.append(out_bytes if op in (P.OP_CONV, P.OP_DWCONV, P.OP_DECONV2) else 0)
    if not rows:
        return None, []
    if os.environ.get("VSE_STEP_TABLE"):       # per-step dump for profiling notes (profiles/)
        with open(os.environ["VSE_STEP_TABLE"], "w") as f:
            f.write("plan kernel ms bytes GB/s GFLOP\n")
            for which, k, t, b, fl in rows:
                f.write(f"{which} {k} {t:.4f} {b} {b / max(t, 1e-6) / 1e6:.1f} {fl / 1e9:.3f}\n")
    by_kernel = {}
    for which, k, t, b, f in rows:
        a = by_kernel.setdefault(k, [0.0, 0, 0, 0])
        a[0] += t; a[1] += b; a[2] += f; a[3] += 1
    top = max(by_kernel.items(), key=lambda kv: kv[1][0])
    return top, sorted(((k, v[0], v[1], v[2], v[3]) for k, v in by_kernel.items()), key=lambda r: -r[1])


def hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            with open(path) as f:
                d = json.load(f)
            for key in ("hbm_gbs", "hbm_gbps", "hbm_gb_s"):
                if key in d:
                    return float(d[key]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def ncu_traffic():
    """DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture, if any."""
    path = os.path.join(ROOT, "profiles", "top_kernel_traffic.json")
    if os.path.exists(path):
        try:
            with open(path) as f:
                return json.load(f)
        except Exception:
            return None
    return None


def run_b200(args, rank, local_rank, world):
    import torch
    from video_subtitle_extractor_b200 import engine as E
    from video_subtitle_extractor_b200 import shard, weights
    from video_subtitle_extractor_b200.synth import SynthStream

    if E.device_count() == 0:
        raise SystemExit("bench.py: no CUDA device (the engine has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = f"cuda:{local_rank}"
    numa = shard.bind_to_gpu_numa_node(local_rank) if world > 1 else None     # before any page-locked allocation
    if world > 1:
        import torch.distributed as dist
        # NCCL's own log lines (version banner, NCCL_DEBUG=INFO) go to stderr: stdout carries the one JSON line only
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device(dev))
    # one-time plan broadcast (rank 0 reads the packed plans; NCCL over NVLink)
    blobs = [weights.load_plan_blob(DET), weights.load_plan_blob(REC)] if rank == 0 else None
    det_blob, rec_blob = shard.broadcast_blobs(blobs, 0, dev) if world > 1 else blobs

    B, H, W = args.batch, args.height, args.width
    stream = SynthStream(H, W)
    video_frames = {}
    if args.video:      # real frames: every frame_stride-th frame of the video, decoded once up front
        video_frames = video_frames_at(args.video, [frame_index(p, k, world * B, args.frame_stride, B) for p in range(args.pool)
                                                    for k in range(world * B)])
        H, W = next(iter(video_frames.values())).shape[:2]
        args.height, args.width = H, W
    # rank r owns frames [r*B, (r+1)*B) of every global batch of world*B frames
    host_batches, dev_batches = [], []
    for p in range(args.pool):
        lo, hi = shard.frame_range(rank, world, world * B)
        idx = [frame_index(p, k, world * B, args.frame_stride, B) for k in range(lo, hi)]
        pinned = torch.empty((len(idx), H, W, 3), dtype=torch.uint8, pin_memory=True)
        arr = pinned.numpy()
        for j, i in enumerate(idx):
            arr[j] = video_frames[i] if args.video else stream.frame(i)
        host_batches.append(pinned)
        dev_batches.append(pinned.to(dev))
    torch.cuda.synchronize()

    # V2 recognisers (ResNet + BiLSTM) read 32-pixel-high crops (reference backend/tools/paddle_model_config.py:94-97)
    rec_h = 32 if REC.startswith("V2/") else 48
    mode = dict(E.bench_mode())
    if args.precision == "mixed":
        mode = dict(E.mixed_mode())
    elif args.precision:
        mode = dict(precision={"fp16": E.PRECISION_FP16, "fp32": E.PRECISION_FP32, "tf32": E.PRECISION_TF32,
                               "fp32_tc": E.PRECISION_FP32_TC}[args.precision])
    mode["flags"] = mode.get("flags", 0) | args.flags
    det_split = bool(mode["flags"] & E.FLAG_DET_FP32_TC) and mode["precision"] == E.PRECISION_FP16
    prec_name = {E.PRECISION_FP16: "fp16 activations, fp32 accumulate (NOT the parity mode: DESIGN.md §5)",
                 E.PRECISION_FP32: "fp32 activations, CUDA-core kernels",
                 E.PRECISION_TF32: "fp32 activations, tf32 tensor-core products",
                 E.PRECISION_FP32_TC: "fp32 activations; tensor-core products on fp16 hi+lo splits of both operands "
                                      "(3 MMAs per product), fp32 accumulate"}[mode["precision"]]
    if det_split:
        prec_name = ("detector: fp32 activations, tensor-core products on fp16 hi+lo splits of both operands (3 MMAs per product); "
                     "recogniser: fp16 activations, fp32 accumulate")
    eng = E.Engine(device=local_rank, rec_image_h=rec_h, **mode)
    eng.load_plan(E.PLAN_DET, det_blob, DET)
    eng.load_plan(E.PLAN_REC, rec_blob, REC)
    hs, ws = [H] * B, [W] * B
    frame_bytes = H * W * 3

    def step(batches, k, mem_kind):
        t = batches[k % len(batches)]
        base = t.data_ptr()
        return eng.run_device([base + j * frame_bytes for j in range(B)], hs, ws, None, mem_kind=mem_kind)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def prefetch(batches, k):
        base = batches[k % len(batches)].data_ptr()
        eng.prefetch([base + j * frame_bytes for j in range(B)], hs, ws, None, mem_kind=E.MEM_PINNED)

    def timed(batches, mem_kind):
        # host-resident batches: the copy of step k+1 is issued (vse_prefetch, copy stream) before step k runs, so every
        # step's host->device copy is inside the timed region but overlaps the previous step's kernels
        pf = mem_kind != E.MEM_DEVICE
        for k in range(args.warmup):
            step(batches, k, mem_kind)
        barrier()
        dev_ms = 0.0
        l0 = eng.launch_count
        t0 = time.perf_counter()
        n_lines = 0
        widths = []
        if pf:
            prefetch(batches, 0)
        for k in range(args.steps):
            if pf and k + 1 < args.steps:
                prefetch(batches, k + 1)
            res = step(batches, k, mem_kind)
            dev_ms += float(eng.last_timings[7])
            n_lines += sum(len(r.quads) for r in res)
            widths += [int(w_) for r in res for w_ in r.rec_widths]
        barrier()
        wall = time.perf_counter() - t0
        timed.mean_width = float(np.mean(widths)) if widths else 0.0
        return wall, dev_ms / 1e3, eng.launch_count - l0, n_lines, res

    # clocks are sampled on rank 0 only (one nvidia-smi child per box, not one per rank, inside the timed region)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    wall, dev_s, launches, n_lines, last = timed(dev_batches, E.MEM_DEVICE)
    mean_width = timed.mean_width
    clocks = sampler.stop() if sampler else None
    wall_e2e, dev_s_e2e, _, _, last_e2e = timed(host_batches, E.MEM_PINNED)
    stage_ms = [float(x) for x in eng.last_timings]
    # one extra un-prefetched call: its stage 0 is the bare host->device copy time of one batch on this box (PCIe rate)
    step(host_batches, 1, E.MEM_PINNED)
    h2d_alone_ms = float(eng.last_timings[0])

    wall_max = shard.max_over_ranks(wall, dev)
    wall_e2e_max = shard.max_over_ranks(wall_e2e, dev)
    per_rank = [(rank, wall / args.steps * 1e3, dev_s / args.steps * 1e3, wall_e2e / args.steps * 1e3, n_lines / max(args.steps, 1), numa)]
    if world > 1:
        bucket = [None] * world
        torch.distributed.all_gather_object(bucket, per_rank[0])
        per_rank = bucket
    total_frames = world * B * args.steps
    value = total_frames / wall_max
    e2e_value = total_frames / wall_e2e_max
    d2h = sum(4 + len(r.quads) * (32 + 4 + 4 + 4 + 4) + sum(len(i) for i in r.ids) * 4 for r in last_e2e)

    out = None
    if rank == 0:
        step(dev_batches, 0, E.MEM_DEVICE)    # make the last run of both plans a full resident batch
        top, table = step_roofline(eng, E)
        peak, peak_src = hbm_peak()
        roofline = None
        if top is not None:
            name, (t_ms, bytes_, flops, n_launch) = top
            achieved = bytes_ / (t_ms * 1e-3) / 1e9
            traffic = ncu_traffic()
            roofline = {"bound": "hbm", "kernel": name, "achieved": achieved, "peak": peak, "unit": "GB/s",
                        "frac": achieved / peak, "traffic": (traffic or {}).get("dram_bytes_per_launch") if args.config == 1 else None,
                        "peak_source": peak_src, "launches_per_step": n_launch,
                        "algorithmic_bytes_per_launch": bytes_ / n_launch, "avg_launch_ms": t_ms / n_launch,
                        "tflops": flops / (t_ms * 1e-3) / 1e12,
                        "per_kernel_ms": {k: round(t, 4) for k, t, _, _, _ in table},
                        "per_kernel_gbs": {k: round(b / max(t, 1e-6) / 1e6, 1) for k, t, b, _, _ in table}}
        if roofline is not None and flops / max(bytes_, 1) > 215.0:
            # the server models (309-649 FLOP/B) sit right of the ridge: tensor-pipe roofline.  ALGORITHMIC
            # flops (one product per multiply-add; the fp32 tensor-core mode issues three MMAs per product) against the
            # measured sustained bf16 rate
            tpeak, tsrc = 1407.5, "fallback (SURVEY.md §8d)"
            try:
                with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                    tpeak, tsrc = float(json.load(f)["bf16_tflops_sustained"]), "measured (MEASURED_PEAKS.json: bf16_tflops_sustained)"
            except Exception:
                pass
            roofline.update({"bound": "tensor", "achieved": roofline["tflops"], "peak": tpeak, "unit": "TFLOP/s",
                             "frac": roofline["tflops"] / tpeak, "peak_source": tsrc, "hbm_gbs": achieved,
                             "executed_tflops": 3 * roofline["tflops"], "executed_frac": 3 * roofline["tflops"] / tpeak,
                             "note": "achieved / frac count ALGORITHMIC FLOPs (one product per multiply-add); the fp32 tensor-core mode "
                                     "issues 3 fp16 MMAs per product (hi*Wh + lo*Wh + hi*Wl): executed_* is the rate the tensor pipe runs at"})
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            oracle, cores = cpu_oracle(det_blob, rec_blob)
            frames = [host_batches[0].numpy()[j] for j in range(B)]
            if args.config != 1:
                frames = frames[:1]      # the server graphs cost ~12 s per frame on the CPU: one frame is the bounded sample
            fps, n, dt = time_cpu(oracle, frames, args.cpu_seconds)
            cpu = {"value": fps, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": f"{n} frames = {n / B:.2f} pass(es) over the {B} frames of one step ({H}x{W}) in {dt:.1f} s: CPU restatement "
                             f"of the reference path (torch-CPU fp32 + cv2), one frame per call, rec batches <= 6"}
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": wall_max / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32+f16" if det_split else "f16" if mode["precision"] == E.PRECISION_FP16 else "f32", "data": "real video frames" if args.video else "synthetic",
            "config": workload_config(args, world),
            "precision": prec_name + (f"; vse_config.flags={args.flags}" if args.flags else ""),
            "text_lines_per_frame": n_lines / max(args.steps * B, 1), "mean_padded_rec_width": mean_width,
            "per_rank": [{"rank": r, "ms_per_step": round(a, 4), "device_ms_per_step": round(b, 4), "e2e_ms_per_step": round(c, 4),
                          "text_lines_per_step": d, "numa_node": nn} for r, a, b, c, d, nn in per_rank],
            "device_ms_per_step": dev_s / args.steps * 1e3, "stage_ms_last_e2e_step": stage_ms,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": B * frame_bytes, "d2h_bytes_per_step": d2h,
                    "ms_per_step": wall_e2e_max / args.steps * 1e3, "device_ms_per_step": dev_s_e2e / args.steps * 1e3,
                    "h2d_ms_per_step_alone": h2d_alone_ms, "h2d_gbs": B * frame_bytes / max(h2d_alone_ms, 1e-6) / 1e6,
                    "overlap": "vse_prefetch: the copy of step k+1 is issued before step k runs (copy stream)"},
            "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
        }
    eng.close()
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()
    if out is not None:
        emit(out)


# --------------------------------------------------------------------------------------------------
# BASELINE configs[3]: four sample videos, fast mode, every video's schedule sharded over the ranks — whole jobs
# --------------------------------------------------------------------------------------------------
VIDEO_JOBS = [("test_en.mp4", "en", "V4/en_rec_fast", 97), ("test_cn.mp4", "ch", "V4/ch_rec_fast", 6625),
              ("test_japan.mp4", "japan", "V3/japan_rec_fast", 4401), ("test_korean.flv", "korean", "V3/korean_rec_fast", 3690)]


def run_videos(args, rank, local_rank, world):
    """One step = the four videos through job.fast_mode_job (decoder threads -> pinned ring -> vse_prefetch / vse_run ->
    raw.txt lines -> gather by frame -> de-dup -> .srt text on every rank).  Recognisers follow the reference's fallback chain
    (backend/tools/paddle_model_config.py:73-82).  value = OCR'd frames of all videos / wall time of the slowest rank."""
    import warnings
    import torch
    from video_subtitle_extractor_b200 import charset, engine as E, job, shard, weights
    warnings.simplefilter("ignore", RuntimeWarning)       # dictionaries of ch / japan / korean are not in this image
    torch.cuda.set_device(local_rank)
    dev = f"cuda:{local_rank}"
    if world > 1:
        shard.bind_to_gpu_numa_node(local_rank)        # decode thread + page-locked ring next to the GPU
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        torch.distributed.init_process_group("nccl", device_id=torch.device(dev))
    vids = os.path.join(ROOT, "tests", "golden", "_videos")
    jobs = [(os.path.join(vids, v), lang, rec, n) for v, lang, rec, n in VIDEO_JOBS if os.path.exists(os.path.join(vids, v)) and weights.have_plan(rec)]
    if not jobs:
        raise SystemExit("bench.py --config 3: tests/golden/_videos/ or the packed recogniser plans are missing")
    eng = E.Engine(device=local_rank, **E.bench_mode())
    eng.load_plan(E.PLAN_DET, weights.load_plan_blob("V4/ch_det_fast"), "V4/ch_det_fast")

    def one_pass(stats):
        n_frames, n_lines, n_subs = 0, 0, 0
        for path, lang, rec, n_cls in jobs:
            t_lp = time.perf_counter()
            eng.load_plan(E.PLAN_REC, weights.load_plan_blob(rec), rec)
            stats["load_plan_s"] = stats.get("load_plan_s", 0.0) + time.perf_counter() - t_lp
            res = job.fast_mode_job(eng, path, charset.characters(lang, None, n_cls), rank=rank, world=world, batch=args.batch,
                                    rec_char_type=lang, stats=stats, write_srt=rank == 0)
            n_frames += res.frames_ocr
            n_lines += len(res.lines)
            n_subs += len(res.subtitles)
        return n_frames, n_lines, n_subs

    for _ in range(args.warmup):
        one_pass({})
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
    stats = {}
    l0 = eng.launch_count
    t0 = time.perf_counter()
    for _ in range(args.steps):
        n_frames, n_lines, n_subs = one_pass(stats)
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
    wall = time.perf_counter() - t0
    wall_max = shard.max_over_ranks(wall, dev)
    mine = (rank, n_frames, wall / args.steps, stats.get("feed_wait_s", 0) / args.steps, stats.get("engine_s", 0) / args.steps,
            stats.get("frames_decoded", 0) // args.steps,
            {k: round(stats.get(k, 0.0) / args.steps, 3) for k in ("load_plan_s", "feed_setup_s", "run_feed_s", "lines_s", "gather_s", "srt_s")})
    per_rank = [mine]
    if world > 1:
        bucket = [None] * world
        torch.distributed.all_gather_object(bucket, mine)
        per_rank = bucket
    total = sum(r[1] for r in per_rank)
    launches = eng.launch_count - l0
    eng.close()
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()
    if rank == 0:
        emit({"metric": METRIC, "value": total * args.steps / wall_max, "unit": UNIT, "n_gpus": world, "steps": args.steps,
              "warmup": args.warmup, "ms_per_step": wall_max / args.steps * 1e3, "higher_is_better": True, "scaling": "strong",
              "vs_baseline": None, "dtype": "f32", "data": "real video frames",
              "config": {"workload": "BASELINE configs[3]: " + ", ".join(os.path.basename(j[0]) for j in jobs) + " in fast mode, whole jobs "
                                     "(decode -> vse_run -> raw.txt -> de-dup -> .srt), each video's schedule sharded over the ranks",
                         "baseline_config": 3, "frames_per_step": total, "raw_lines": n_lines, "subtitles": n_subs,
                         "models": ["V4/ch_det_fast"] + [j[2] for j in jobs], "frames_per_vse_run": args.batch},
              "e2e": {"value": total * args.steps / wall_max, "unit": UNIT, "note": "the job IS end to end: frames come from the "
                      "video decoder through pinned host buffers, results go back to host text"},
              "per_rank": [{"rank": r, "frames_ocr": n, "s_per_step": round(w, 3), "feed_wait_s": round(fw, 3), "engine_s": round(es, 3),
                            "frames_decoded": fd, "phases_s": ph} for r, n, w, fw, es, fd, ph in per_rank],
              "limiter": f"video decode (cv2 / ffmpeg, {job.default_decoders(world)} decoder thread(s) per rank)" if mine[3] > mine[4] else "engine",
              "gpu_launches": launches})


_JSON_FD = None


def claim_stdout():
    """stdout carries the one JSON line and nothing else: libraries that print to file descriptor 1 (NCCL's version banner
    does, whatever NCCL_DEBUG_FILE says) are sent to stderr, and emit() writes the result to the real stdout."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit(obj):
    line = (json.dumps(obj) + "\n").encode()
    sys.stdout.flush()
    if _JSON_FD is None:
        os.write(1, line)
    else:
        os.write(_JSON_FD, line)


def main():
    args = parse_args()
    apply_model_args(args)
    rank, local_rank, world = env_rank()
    if not (args.impl != "reference" and world == 1 and args.gpus > 1):   # the torchrun re-launch keeps the child's stdout
        claim_stdout()
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world == 1 and args.gpus > 1:
        # convenience: re-launch under torchrun (the driver launches torchrun itself)
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29511", os.path.abspath(__file__)] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    if args.config == 3:
        run_videos(args, rank, local_rank, world)
        return
    run_b200(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
