"""ORACLE-SIDE CHECKER (test infrastructure, never shipped on the product path).

Executes a compiled ``Plan`` (video_subtitle_extractor_b200/plan.py) step by step on
torch-CPU fp32.  Purpose: prove on a machine without a GPU that the plan compiler's
fusion / BN folding / concat aliasing preserves the arithmetic of the shipped graph
(compare with oracle/graph_interp.py), and give the GPU tests a per-step reference
for the CUDA runtime (csrc/runtime.cu executes the very same step list).
"""
from __future__ import annotations

from typing import Dict, List, Optional

import numpy as np
import torch
import torch.nn.functional as F

from video_subtitle_extractor_b200 import plan as P


def _act(x: torch.Tensor, act: int, slope: float = 0.0, offset: float = 0.0) -> torch.Tensor:
    if act == P.ACT_NONE:
        return x
    if act == P.ACT_RELU:
        return F.relu(x)
    if act == P.ACT_RELU6:
        return torch.clamp(x, 0.0, 6.0)
    if act == P.ACT_HSWISH:
        return x * torch.clamp(x + 3.0, 0.0, 6.0) / 6.0
    if act == P.ACT_HSIGMOID:
        return torch.clamp(x * slope + offset, 0.0, 1.0)
    if act == P.ACT_SWISH:
        return x * torch.sigmoid(x)
    if act == P.ACT_SIGMOID:
        return torch.sigmoid(x)
    raise ValueError(act)


class PlanInterpreter:
    """Values are kept as NCHW float tensors ([N,C] for pooled vectors)."""

    def __init__(self, plan: P.Plan):
        self.plan = plan
        # weights converted once to the layouts torch wants (keeps the CPU baseline honest: no per-call re-layout)
        self._wcache: Dict[int, Dict[str, torch.Tensor]] = {}
        for k, s in enumerate(plan.steps):
            c: Dict[str, torch.Tensor] = {}
            for name, arr in s.w.items():
                c[name] = torch.from_numpy(np.array(arr, dtype=np.float32, copy=True))
            if s.op in (P.OP_CONV, P.OP_STEM):
                c["weight"] = c["weight"].permute(0, 3, 1, 2).contiguous()
            elif s.op == P.OP_DWCONV:
                c["weight"] = c["weight"].permute(2, 0, 1).unsqueeze(1).contiguous()
            elif s.op == P.OP_DECONV2:
                c["weight"] = c["weight"].permute(3, 2, 0, 1).contiguous()
            elif s.op == P.OP_VECLIN:
                c["weight"] = c["weight"].t().contiguous()
            self._wcache[id(s)] = c

    def _w(self, s: P.Step, name: str) -> torch.Tensor:
        return self._wcache[id(s)][name]

    def _epilogue(self, s: P.Step, y: torch.Tensor, env) -> torch.Tensor:
        shp = [1, -1, 1, 1] if y.dim() == 4 else [1, -1]
        y = y + self._w(s, "bias").reshape(shp)
        y = _act(y, s.p["act"], s.p.get("hs_slope", 0.0), s.p.get("hs_offset", 0.0))
        if s.p.get("has_post"):
            y = y * self._w(s, "post_scale").reshape(shp) + self._w(s, "post_shift").reshape(shp)
        if s.p.get("has_res"):
            y = y + env[s.ins[1]]
        return _act(y, s.p.get("act2", P.ACT_NONE))

    @torch.no_grad()
    def run(self, x: torch.Tensor, valid_w: Optional[List[int]] = None, keep_all: bool = False, store_hook=None):
        """x: normalised float NCHW, or uint8 NHWC (normalised here with the plan's constants;
        columns >= valid_w[n] are zero in normalised space, as the reference's right zero-pad).
        store_hook(step_index, step, tensor) -> tensor, if given, is applied to every step output before it is stored
        (tools/precision_study.py emulates the engine's 16-bit activation storage with it)."""
        plan = self.plan
        if x.dtype == torch.uint8:
            sc = torch.tensor(plan.norm_scale).reshape(1, 1, 1, 3)
            sh = torch.tensor(plan.norm_shift).reshape(1, 1, 1, 3)
            xf = x.float() * sc + sh
            if valid_w is not None:
                for n, vw in enumerate(valid_w):
                    xf[n, :, vw:, :] = 0.0
            x = xf.permute(0, 3, 1, 2).contiguous()
        env: Dict[int, torch.Tensor] = {plan.input_vid: x.float()}
        views: Dict[int, List] = {}
        vals = plan.values

        def put(vid: int, t: torch.Tensor):
            if store_hook is not None:
                t = store_hook(step_no[0], cur[0], t)
            env[vid] = t
            v = vals[vid]
            if v.alias_of >= 0:
                views.setdefault(v.alias_of, []).append((v.alias_coff, t))

        def materialise(vid: int):
            if vid in env or vid not in views:
                return
            parts = sorted(views[vid], key=lambda p: p[0])
            env[vid] = torch.cat([t for _, t in parts], dim=1)

        step_no, cur = [0], [None]
        for k, s in enumerate(plan.steps):
            step_no[0], cur[0] = k, s
            for v in s.ins:
                materialise(v)
            op = s.op
            if op in (P.OP_CONV, P.OP_STEM):
                w = self._w(s, "weight")
                y = F.conv2d(env[s.ins[0]], w, None, (s.p["sh"], s.p["sw"]), (s.p["ph"], s.p["pw"]))
                put(s.out, self._epilogue(s, y, env))
            elif op == P.OP_DWCONV:
                w = self._w(s, "weight")
                y = F.conv2d(env[s.ins[0]], w, None, (s.p["sh"], s.p["sw"]), (s.p["ph"], s.p["pw"]), 1, w.shape[0])
                put(s.out, self._epilogue(s, y, env))
            elif op == P.OP_DECONV2:
                w = self._w(s, "weight")  # [cin,cout,kh,kw]
                y = F.conv_transpose2d(env[s.ins[0]], w, None, 2)
                put(s.out, self._epilogue(s, y, env))
            elif op == P.OP_VECLIN:
                y = env[s.ins[0]] @ self._w(s, "weight")
                put(s.out, self._epilogue(s, y, env))
            elif op == P.OP_GPOOL:
                put(s.out, env[s.ins[0]].mean(dim=(2, 3)))
            elif op == P.OP_CHSCALE:
                xx = env[s.ins[0]]
                y = xx * env[s.ins[1]][:, :, None, None]
                put(s.out, xx + y if s.p.get("residual") else y)
            elif op == P.OP_POOL:
                k, st, pd = (s.p["kh"], s.p["kw"]), (s.p["sh"], s.p["sw"]), (s.p["ph"], s.p["pw"])
                if s.p["is_max"]:
                    y = F.max_pool2d(env[s.ins[0]], k, st, pd, ceil_mode=bool(s.p["ceil"]))
                else:
                    y = F.avg_pool2d(env[s.ins[0]], k, st, pd, ceil_mode=bool(s.p["ceil"]),
                                     count_include_pad=not s.p["exclusive"])
                put(s.out, y)
            elif op == P.OP_UPSAMPLE:
                sc = s.p["scale"]
                y = env[s.ins[0]].repeat_interleave(sc, dim=2).repeat_interleave(sc, dim=3)
                if s.p.get("has_add"):
                    y = env[s.ins[1]] + y
                put(s.out, y)
            elif op == P.OP_ADD:
                put(s.out, _act(env[s.ins[0]] + env[s.ins[1]], s.p.get("act", P.ACT_NONE)))
            elif op == P.OP_COPY:
                views.setdefault(s.out, []).append((s.p["coff"], env[s.ins[0]]))
            elif op == P.OP_ELTWISE:
                xx = env[s.ins[0]]
                shp = [1, -1, 1, 1] if xx.dim() == 4 else [1, -1]
                y = xx * self._w(s, "scale").reshape(shp) + self._w(s, "shift").reshape(shp)
                put(s.out, _act(y, s.p["act"], s.p.get("hs_slope", 0.0), s.p.get("hs_offset", 0.0)))
            elif op == P.OP_LAYERNORM:
                xx = env[s.ins[0]]  # [B,C,1,T]
                y = F.layer_norm(xx.permute(0, 2, 3, 1), [xx.shape[1]], self._w(s, "gamma"), self._w(s, "beta"), s.p["eps"])
                put(s.out, y.permute(0, 3, 1, 2).contiguous())
            elif op == P.OP_ATTN:
                qkv = env[s.ins[0]]  # [B, 3*H*D, 1, T]
                B, _, _, T = qkv.shape
                H, D = s.p["heads"], s.p["dim"]
                t = qkv[:, :, 0, :].permute(0, 2, 1).reshape(B, T, 3, H, D).permute(2, 0, 3, 1, 4)
                q, k, v = t[0] * s.p["qscale"], t[1], t[2]
                pr = torch.softmax(q @ k.transpose(-1, -2), dim=-1)
                ctx = (pr @ v).permute(0, 2, 1, 3).reshape(B, T, H * D)
                put(s.out, ctx.permute(0, 2, 1).unsqueeze(2).contiguous())
            elif op == P.OP_SOFTMAX:
                put(s.out, torch.softmax(env[s.ins[0]], dim=1))
            elif op == P.OP_LSTM:
                put(s.out, self._lstm(s, env[s.ins[0]]))
            else:
                raise NotImplementedError(P.OP_NAMES[op])
        for v in plan.output_vids:
            materialise(v)
        outs = [env[v] for v in plan.output_vids]
        if keep_all:
            for v in list(views):
                materialise(v)
            return outs, env
        return outs

    def _lstm(self, s: P.Step, x: torch.Tensor) -> torch.Tensor:
        """Recurrent half of one (bi)LSTM layer.  ``x`` [B, ndir*4*hidden, 1, T] holds the gate pre-activations
        W_ih x_t + b_ih + b_hh of every direction (the preceding CONV step); gate order i,f,g,o as in Paddle's ``rnn`` op
        (oracle/graph_interp.py::_rnn follows the shipped V2/ch_rec graph, op#140).  The whole padded width is the sequence,
        as in the reference (the recogniser pads each batch before the network)."""
        hidden, ndir = s.p["hidden"], s.p["ndir"]
        w_hh = self._w(s, "weight").reshape(ndir, 4 * hidden, hidden)
        seq = x[:, :, 0, :].permute(2, 0, 1)  # [T,B,ndir*4H]
        T, B, _ = seq.shape
        outs = []
        for d in range(ndir):
            xs = seq[:, :, d * 4 * hidden:(d + 1) * 4 * hidden]
            h = torch.zeros(B, hidden)
            c = torch.zeros(B, hidden)
            hs = [None] * T
            for t in (range(T) if d == 0 else range(T - 1, -1, -1)):
                g = xs[t] + h @ w_hh[d].t()
                i, f, gg, o = g.chunk(4, dim=1)
                c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
                h = torch.sigmoid(o) * torch.tanh(c)
                hs[t] = h
            outs.append(torch.stack(hs, 0))
        seq = torch.cat(outs, dim=2)
        return seq.permute(1, 2, 0).unsqueeze(2).contiguous()
