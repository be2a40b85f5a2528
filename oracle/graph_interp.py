"""ORACLE (test infrastructure, never shipped on the product path).

CPU restatement of what ``paddle.inference`` does with the reference's shipped
graphs: executes ``inference.pdmodel`` op by op in fp32 on torch-CPU, with
parameters from ``inference.pdiparams``.  This is the arithmetic behind

* reference backend/tools/subtitle_detect.py:22-26 (``TextDetector`` predictor run)
* reference backend/tools/ocr.py:27 (``PaddleOCR.__call__`` det + rec predictor runs)

paddlepaddle==3.0.0 (README_en.md:186,225) is a pip dependency that is absent
from /root/reference and from this image, so the op semantics are restated from
the published Paddle operator definitions.  The graphs and
weights are the reference's own files; nothing here is copied source.

Parity status: the reference holds no golden vectors for this path;
the anchors that pin this oracle are the observed decodes recorded in SURVEY.md
Appendix E (see tests/test_oracle_cpu.py).  Paddle's own fp32 results may differ in
the last bits (oneDNN summation order, conv+BN folding).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F

from video_subtitle_extractor_b200.loader import Model, Op


def _bcast_y(x: torch.Tensor, y: torch.Tensor, axis: int) -> torch.Tensor:
    """Paddle elementwise broadcasting: align ``y``'s dims to ``x`` starting at ``axis``."""
    if x.dim() == y.dim():
        return y
    if y.dim() > x.dim():
        return y  # handled by caller swapping (scalar-X case)
    if axis == -1:
        axis = x.dim() - y.dim()
    shape = [1] * axis + list(y.shape) + [1] * (x.dim() - axis - y.dim())
    return y.reshape(shape)


def _pool2d(x: torch.Tensor, a: Dict) -> torch.Tensor:
    ptype = a["pooling_type"]
    ksize = list(a["ksize"])
    if a.get("adaptive", False):
        if ksize != [1, 1]:
            raise NotImplementedError("adaptive pool other than 1x1")
        return x.mean(dim=(2, 3), keepdim=True) if ptype == "avg" else x.amax(dim=(2, 3), keepdim=True)
    if a.get("global_pooling", False):
        return x.mean(dim=(2, 3), keepdim=True) if ptype == "avg" else x.amax(dim=(2, 3), keepdim=True)
    strides = list(a["strides"])
    pads = list(a["paddings"])
    if len(pads) == 4:
        if pads[0] != pads[1] or pads[2] != pads[3]:
            raise NotImplementedError("asymmetric pool padding")
        pads = [pads[0], pads[2]]
    ceil_mode = bool(a.get("ceil_mode", False))
    if ptype == "max":
        return F.max_pool2d(x, ksize, strides, pads, ceil_mode=ceil_mode)
    exclusive = bool(a.get("exclusive", True))
    return F.avg_pool2d(x, ksize, strides, pads, ceil_mode=ceil_mode, count_include_pad=not exclusive)


class GraphInterpreter:
    """Executes one shipped model. ``run(x)`` returns the fetch tensors in column order."""

    def __init__(self, model: Model):
        self.model = model
        self.params = {k: torch.from_numpy(np.array(v)) for k, v in model.params.items()}
        self.ops: List[Op] = model.program.ops
        self.feed = model.program.feed_names[0]
        self.fetches = model.program.fetch_names

    # -- helpers -------------------------------------------------------- #
    def _get(self, env: Dict[str, torch.Tensor], name: str) -> torch.Tensor:
        if name in env:
            return env[name]
        return self.params[name]

    @torch.no_grad()
    def run(self, x, record: Optional[Sequence[str]] = None, keep_all: bool = False):
        if isinstance(x, np.ndarray):
            x = torch.from_numpy(np.ascontiguousarray(x))
        env: Dict[str, torch.Tensor] = {self.feed: x.float()}
        for op in self.ops:
            self._exec(op, env)
        outs = [env[n] for n in self.fetches]
        if keep_all:
            return outs, env
        if record:
            return outs, {n: env[n] for n in record}
        return outs

    def _exec(self, op: Op, env: Dict[str, torch.Tensor]) -> None:
        t = op.type
        a = op.attrs
        g = lambda k, i=0: self._get(env, op.inputs[k][i])
        if t in ("feed", "fetch"):
            return
        if t in ("conv2d", "depthwise_conv2d"):
            assert a.get("data_format", "NCHW") in ("NCHW", "AnyLayout")
            pads = list(a["paddings"])
            algo = a.get("padding_algorithm", "EXPLICIT")
            xin = g("Input")
            w = g("Filter")
            if algo == "SAME":
                raise NotImplementedError("SAME padding")
            if algo == "VALID":
                pads = [0, 0]
            if len(pads) == 4:
                if pads[0] != pads[1] or pads[2] != pads[3]:
                    xin = F.pad(xin, (pads[2], pads[3], pads[0], pads[1]))
                    pads = [0, 0]
                else:
                    pads = [pads[0], pads[2]]
            env[op.out("Output")] = F.conv2d(xin, w, None, list(a["strides"]), pads, list(a["dilations"]), a["groups"])
        elif t == "conv2d_transpose":
            pads = list(a["paddings"])
            if len(pads) == 4:
                pads = [pads[0], pads[2]]
            out_pad = a.get("output_padding", []) or [0, 0]
            env[op.out("Output")] = F.conv_transpose2d(
                g("Input"), g("Filter"), None, list(a["strides"]), pads, list(out_pad), a.get("groups", 1),
                list(a["dilations"]))
        elif t == "batch_norm":
            env[op.out("Y")] = F.batch_norm(g("X"), g("Mean"), g("Variance"), g("Scale"), g("Bias"), False, 0.0,
                                            a["epsilon"])
        elif t == "layer_norm":
            xin = g("X")
            bna = a["begin_norm_axis"]
            shape = list(xin.shape[bna:])
            scale = g("Scale") if op.inputs.get("Scale") else None
            bias = g("Bias") if op.inputs.get("Bias") else None
            env[op.out("Y")] = F.layer_norm(xin, shape, scale.reshape(shape) if scale is not None else None,
                                            bias.reshape(shape) if bias is not None else None, a["epsilon"])
        elif t in ("elementwise_add", "elementwise_mul", "elementwise_sub", "elementwise_div"):
            xx, yy = g("X"), g("Y")
            axis = a.get("axis", -1)
            if xx.dim() >= yy.dim():
                yy = _bcast_y(xx, yy, axis)
            else:
                xx = _bcast_y(yy, xx, axis)
            if t == "elementwise_add":
                r = xx + yy
            elif t == "elementwise_mul":
                r = xx * yy
            elif t == "elementwise_sub":
                r = xx - yy
            else:
                r = xx / yy
            env[op.out("Out")] = r
        elif t == "relu":
            env[op.out("Out")] = F.relu(g("X"))
        elif t == "relu6":
            env[op.out("Out")] = torch.clamp(g("X"), 0.0, 6.0)
        elif t == "hard_swish":
            xin = g("X")
            env[op.out("Out")] = xin * torch.clamp(xin + a["offset"], 0.0, a["threshold"]) / a["scale"]
        elif t == "hard_sigmoid":
            env[op.out("Out")] = torch.clamp(g("X") * a["slope"] + a["offset"], 0.0, 1.0)
        elif t == "swish":
            xin = g("X")
            env[op.out("Out")] = xin * torch.sigmoid(a.get("beta", 1.0) * xin)
        elif t == "sigmoid":
            env[op.out("Out")] = torch.sigmoid(g("X"))
        elif t == "pool2d":
            env[op.out("Out")] = _pool2d(g("X"), a)
        elif t in ("nearest_interp_v2", "bilinear_interp_v2"):
            xin = g("X")
            scale = list(a.get("scale", []))
            oh, ow = a.get("out_h", -1), a.get("out_w", -1)
            if scale:
                if len(scale) == 1:
                    scale = scale * 2
                oh, ow = int(xin.shape[2] * scale[0]), int(xin.shape[3] * scale[1])
            if t == "nearest_interp_v2":
                assert not a.get("align_corners", False)
                iy = torch.div(torch.arange(oh) * xin.shape[2], oh, rounding_mode="floor")
                ix = torch.div(torch.arange(ow) * xin.shape[3], ow, rounding_mode="floor")
                env[op.out("Out")] = xin[:, :, iy][:, :, :, ix]
            else:
                env[op.out("Out")] = F.interpolate(xin, (oh, ow), mode="bilinear",
                                                   align_corners=a.get("align_corners", False))
        elif t == "concat":
            env[op.out("Out")] = torch.cat([self._get(env, n) for n in op.inputs["X"]], dim=a["axis"])
        elif t in ("matmul_v2", "matmul"):
            xx, yy = g("X"), g("Y")
            tx = a.get("trans_x", a.get("transpose_X", False))
            ty = a.get("trans_y", a.get("transpose_Y", False))
            if tx:
                xx = xx.transpose(-1, -2)
            if ty:
                yy = yy.transpose(-1, -2)
            r = torch.matmul(xx, yy)
            if t == "matmul" and a.get("alpha", 1.0) != 1.0:
                r = r * a["alpha"]
            env[op.out("Out")] = r
        elif t == "mul":
            xx, yy = g("X"), g("Y")
            xn = a.get("x_num_col_dims", 1)
            lead = list(xx.shape[:xn])
            r = xx.reshape(int(np.prod(lead)), -1) @ yy.reshape(yy.shape[0], -1)
            env[op.out("Out")] = r.reshape(lead + [r.shape[-1]])
        elif t == "scale":
            xin = g("X")
            s, b = a.get("scale", 1.0), a.get("bias", 0.0)
            if a.get("bias_after_scale", True):
                env[op.out("Out")] = xin * s + b if (s != 1.0 or b != 0.0) else xin
            else:
                env[op.out("Out")] = (xin + b) * s
        elif t == "softmax":
            env[op.out("Out")] = torch.softmax(g("X"), dim=a.get("axis", -1))
        elif t == "transpose2":
            env[op.out("Out")] = g("X").permute(list(a["axis"])).contiguous()
        elif t == "reshape2":
            xin = g("X")
            if op.inputs.get("ShapeTensor"):
                shape = [int(self._get(env, n).reshape(-1)[0]) for n in op.inputs["ShapeTensor"]]
            elif op.inputs.get("Shape"):
                shape = [int(v) for v in g("Shape").reshape(-1)]
            else:
                shape = list(a["shape"])
            shape = [xin.shape[i] if s == 0 else s for i, s in enumerate(shape)]
            env[op.out("Out")] = xin.reshape(shape)
        elif t == "flatten_contiguous_range":
            xin = g("X")
            s, e = a["start_axis"], a["stop_axis"]
            if e < 0:
                e += xin.dim()
            env[op.out("Out")] = xin.flatten(s, e)
        elif t == "squeeze2":
            xin = g("X")
            axes = sorted([ax if ax >= 0 else ax + xin.dim() for ax in a["axes"]], reverse=True)
            for ax in axes:
                if xin.shape[ax] == 1:
                    xin = xin.squeeze(ax)
            env[op.out("Out")] = xin
        elif t == "unsqueeze2":
            xin = g("X")
            for ax in sorted(a["axes"]):
                xin = xin.unsqueeze(ax)
            env[op.out("Out")] = xin
        elif t == "slice":
            xin = g("Input")
            axes, starts, ends = a["axes"], a["starts"], a["ends"]
            idx = [slice(None)] * xin.dim()
            for ax, s, e in zip(axes, starts, ends):
                n = xin.shape[ax]
                s = s + n if s < 0 else s
                e = e + n if e < 0 else e
                idx[ax] = slice(min(s, n), min(e, n))
            r = xin[tuple(idx)]
            for ax in sorted(a.get("decrease_axis", []), reverse=True):
                r = r.squeeze(ax)
            env[op.out("Out")] = r
        elif t == "shape":
            env[op.out("Out")] = torch.tensor(list(g("Input").shape), dtype=torch.int32)
        elif t == "fill_constant":
            dt = {2: torch.int32, 3: torch.int64, 5: torch.float32, 0: torch.bool}[a["dtype"]]
            val = a.get("str_value", "") or a["value"]
            shape = list(a["shape"]) or [1]
            env[op.out("Out")] = torch.full(shape, float(val), dtype=torch.float64).to(dt)
        elif t == "fill_constant_batch_size_like":
            ref = g("Input")
            shape = list(a["shape"])
            shape[a.get("output_dim_idx", 0)] = ref.shape[a.get("input_dim_idx", 0)]
            dt = {2: torch.int32, 3: torch.int64, 5: torch.float32}[a["dtype"]]
            env[op.out("Out")] = torch.full(shape, float(a["value"]), dtype=dt)
        elif t in ("assign", "dropout"):
            env[op.out("Out")] = g("X")
        elif t == "cast":
            dt = {2: torch.int32, 3: torch.int64, 5: torch.float32, 0: torch.bool}[a["out_dtype"]]
            env[op.out("Out")] = g("X").to(dt)
        elif t == "rnn":
            env[op.out("Out")] = self._rnn(op, env)
        else:
            raise NotImplementedError(f"op {t}")

    def _rnn(self, op: Op, env: Dict[str, torch.Tensor]) -> torch.Tensor:
        """Paddle ``rnn`` op, mode=LSTM (V2/ch_rec op#140). Gate order i,f,g,o as torch."""
        a = op.attrs
        assert a["mode"] == "LSTM"
        x = self._get(env, op.inp("Input"))  # [T, B, I] time-major
        wl = [self._get(env, n) for n in op.inputs["WeightList"]]
        layers, bidir, hidden = a["num_layers"], a["is_bidirec"], a["hidden_size"]
        ndir = 2 if bidir else 1
        nw = layers * ndir
        weights = wl[:2 * nw]
        biases = wl[2 * nw:]
        out = x
        for l in range(layers):
            dirs = []
            for d in range(ndir):
                k = l * ndir + d
                w_ih, w_hh = weights[2 * k], weights[2 * k + 1]
                b_ih, b_hh = biases[2 * k], biases[2 * k + 1]
                seq = out.flip(0) if d == 1 else out
                T, B, _ = seq.shape
                h = torch.zeros(B, hidden)
                c = torch.zeros(B, hidden)
                xs = seq @ w_ih.t() + b_ih + b_hh
                hs = []
                for ti in range(T):
                    gates = xs[ti] + h @ w_hh.t()
                    i, f, gg, o = gates.chunk(4, dim=1)
                    c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
                    h = torch.sigmoid(o) * torch.tanh(c)
                    hs.append(h)
                hseq = torch.stack(hs, 0)
                if d == 1:
                    hseq = hseq.flip(0)
                dirs.append(hseq)
            out = torch.cat(dirs, dim=2)
        return out
