"""Plan compiler: shipped Paddle graph -> flat list of fused NHWC steps + one weight blob.

The reference hands ``inference.pdmodel``/``.pdiparams`` to ``paddle.inference``
(reference backend/tools/ocr.py:91-113, backend/tools/subtitle_detect.py:16-22), which
interprets ~300 fine-grained NCHW ops per model.  The B200 engine instead runs a short
list of fused steps over channel-last ("pixel-major", [pixels, channels]) activations:

* ``conv2d -> bias -> [learnable-affine] -> act -> [learnable-affine] -> [+residual] -> [act]``
  becomes ONE conv step (scale folded into the weights, the rest in the epilogue);
* ``batch_norm`` is folded into the producing conv;
* ``concat`` disappears: producers write straight into channel slices of the concat buffer;
* ``nearest_interp -> add`` (FPN top-down) is one step;
* the 25-op SVTR attention sub-graph is one ATTN step; reshape/transpose/flatten/squeeze
  are layout bookkeeping only (NHWC with H==1 *is* [B,T,C]).

Everything is derived from the graph (no per-architecture tables), so all 21 shipped
models go through the same path.  The C++ runtime
(csrc/runtime.cu) parses the binary produced by ``Plan.serialize`` and infers run-time
shapes itself; this module never sees image sizes.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional, Tuple

import numpy as np

from .loader import Model, Op

# ---- op codes (mirrored in csrc/plan.h) ------------------------------------ #
OP_CONV, OP_DWCONV, OP_DECONV2, OP_GPOOL, OP_VECLIN, OP_CHSCALE, OP_POOL, OP_UPSAMPLE, OP_ADD, OP_COPY, \
    OP_LAYERNORM, OP_ATTN, OP_ELTWISE, OP_SOFTMAX, OP_STEM, OP_LSTM = range(16)
OP_NAMES = ["CONV", "DWCONV", "DECONV2", "GPOOL", "VECLIN", "CHSCALE", "POOL", "UPSAMPLE", "ADD", "COPY",
            "LAYERNORM", "ATTN", "ELTWISE", "SOFTMAX", "STEM", "LSTM"]

ACT_NONE, ACT_RELU, ACT_HSWISH, ACT_HSIGMOID, ACT_SWISH, ACT_SIGMOID, ACT_RELU6 = range(7)
_ACT_OF = {"relu": ACT_RELU, "hard_swish": ACT_HSWISH, "hard_sigmoid": ACT_HSIGMOID, "swish": ACT_SWISH,
           "sigmoid": ACT_SIGMOID, "relu6": ACT_RELU6}

KIND_IMG, KIND_VEC = 0, 1
DT_ACT, DT_F32, DT_U8 = 0, 1, 2

CPAD = 8  # channel stride granularity (8 halves = 16 B: vector / TMA alignment)


def pad_c(c: int) -> int:
    return (c + CPAD - 1) // CPAD * CPAD


# --------------------------------------------------------------------------- #
# IR
# --------------------------------------------------------------------------- #

@dataclass
class Value:
    vid: int
    channels: int
    kind: int = KIND_IMG
    dtype: int = DT_ACT
    alias_of: int = -1      # view into another value's buffer (concat slices)
    alias_coff: int = 0
    name: str = ""

    @property
    def cstride(self) -> int:
        return pad_c(self.channels) if self.kind == KIND_IMG else self.channels


@dataclass
class Step:
    op: int
    ins: List[int]
    out: int
    p: Dict[str, Any] = field(default_factory=dict)       # int/float params
    w: Dict[str, np.ndarray] = field(default_factory=dict)  # named fp32 arrays
    src_ops: List[int] = field(default_factory=list)      # graph op indices covered (debug)

    def __repr__(self):
        ws = {k: list(v.shape) for k, v in self.w.items()}
        return f"{OP_NAMES[self.op]} {self.ins}->{self.out} {self.p} {ws}"


@dataclass
class _Sym:
    """Symbolic handle for a graph variable during lowering."""
    kind: str                 # 'img' | 'vec' | 'const' | 'qkv' | 'qkv_t' | 'q' | 'k' | 'v' | 'kT' | 'scores' | 'ctx' | 'ctx_t'
    vid: int = -1
    tag: str = "nchw"         # logical view of an H==1 image: nchw | bcw | btc | b1tc | tbc
    const: Optional[np.ndarray] = None
    extra: Dict[str, Any] = field(default_factory=dict)


class PlanError(RuntimeError):
    pass


# --------------------------------------------------------------------------- #
# lowering
# --------------------------------------------------------------------------- #

class _Lowerer:
    def __init__(self, model: Model):
        self.model = model
        self.P = model.params
        self.values: List[Value] = []
        self.nodes: List[Step] = []
        self.sym: Dict[str, _Sym] = {}

    def new_value(self, channels: int, kind: int = KIND_IMG, dtype: int = DT_ACT, name: str = "") -> int:
        v = Value(len(self.values), int(channels), kind, dtype, name=name)
        self.values.append(v)
        return v.vid

    def get(self, name: str) -> _Sym:
        if name in self.sym:
            return self.sym[name]
        if name in self.P:
            return _Sym("const", const=self.P[name])
        raise PlanError(f"unknown variable {name}")

    def emit(self, op: int, ins: List[int], out: int, opi: int, p=None, w=None) -> Step:
        s = Step(op, list(ins), out, dict(p or {}), dict(w or {}), [opi])
        self.nodes.append(s)
        return s

    # -- per-op lowering --------------------------------------------------- #
    def run(self):
        ops = self.model.program.ops
        for i, op in enumerate(ops):
            fn = getattr(self, "_op_" + op.type, None)
            if fn is None:
                raise PlanError(f"op #{i} {op.type}: no lowering")
            fn(i, op)
        return self

    def _op_feed(self, i, op: Op):
        vid = self.new_value(3, KIND_IMG, DT_U8, name="input_u8")
        self.sym[op.out("Out")] = _Sym("img", vid)
        self.input_vid = vid

    def _op_fetch(self, i, op: Op):
        s = self.get(op.inp("X"))
        if not hasattr(self, "fetch_vids"):
            self.fetch_vids = {}
        self.fetch_vids[op.attrs.get("col", 0)] = s.vid

    def _conv_common(self, i, op: Op, depthwise_hint: bool):
        a = op.attrs
        x = self.get(op.inp("Input"))
        w = self.P[op.inp("Filter")]
        cout, cin_g, kh, kw = w.shape
        groups = a["groups"]
        pads = list(a["paddings"])
        if a.get("padding_algorithm", "EXPLICIT") == "VALID":
            pads = [0, 0]
        if len(pads) == 4:
            if pads[0] != pads[1] or pads[2] != pads[3]:
                raise PlanError("asymmetric conv padding")
            pads = [pads[0], pads[2]]
        if list(a["dilations"]) != [1, 1]:
            raise PlanError("dilated conv")
        sh, sw = a["strides"]
        if x.kind == "vec":
            if (kh, kw) != (1, 1) or groups != 1:
                raise PlanError("non-1x1 conv on pooled vector")
            vid = self.new_value(cout, KIND_VEC, DT_F32)
            self.emit(OP_VECLIN, [x.vid], vid, i, p=dict(cin=cin_g, cout=cout),
                      w=dict(weight=w.reshape(cout, cin_g).astype(np.float32)))
            self.sym[op.out("Output")] = _Sym("vec", vid)
            return
        if x.kind != "img":
            raise PlanError(f"conv input kind {x.kind}")
        cin = self.values[x.vid].channels
        vid = self.new_value(cout)
        p = dict(kh=kh, kw=kw, sh=sh, sw=sw, ph=pads[0], pw=pads[1], cin=cin, cout=cout)
        if groups == 1:
            if cin_g != cin:
                raise PlanError("conv channel mismatch")
            opc = OP_STEM if self.values[x.vid].dtype == DT_U8 else OP_CONV
            # canonical weight layout: [cout][kh][kw][cin]
            self.emit(opc, [x.vid], vid, i, p=p, w=dict(weight=np.ascontiguousarray(w.transpose(0, 2, 3, 1))))
        elif groups == cin and cin_g == 1 and cout == cin:
            # depthwise: [kh][kw][c]
            self.emit(OP_DWCONV, [x.vid], vid, i, p=p, w=dict(weight=np.ascontiguousarray(w[:, 0].transpose(1, 2, 0))))
        else:
            raise PlanError(f"grouped conv groups={groups} cin={cin} cout={cout}")
        self.sym[op.out("Output")] = _Sym("img", vid)

    def _op_conv2d(self, i, op):
        self._conv_common(i, op, False)

    def _op_depthwise_conv2d(self, i, op):
        self._conv_common(i, op, True)

    def _op_conv2d_transpose(self, i, op: Op):
        a = op.attrs
        x = self.get(op.inp("Input"))
        w = self.P[op.inp("Filter")]  # [cin, cout, kh, kw]
        cin, cout, kh, kw = w.shape
        if (kh, kw) != (2, 2) or list(a["strides"]) != [2, 2] or any(a["paddings"]) or a.get("groups", 1) != 1:
            raise PlanError("conv2d_transpose other than 2x2 s2 p0")
        vid = self.new_value(cout)
        # canonical: [kh][kw][cout][cin]
        self.emit(OP_DECONV2, [x.vid], vid, i, p=dict(cin=cin, cout=cout),
                  w=dict(weight=np.ascontiguousarray(w.transpose(2, 3, 1, 0))))
        self.sym[op.out("Output")] = _Sym("img", vid)

    def _affine(self, i, x: _Sym, scale: np.ndarray, shift: np.ndarray, out_name: str):
        """Per-channel affine on an img/vec value."""
        c = self.values[x.vid].channels
        scale = np.broadcast_to(np.asarray(scale, np.float32).reshape(-1), (c,)).copy()
        shift = np.broadcast_to(np.asarray(shift, np.float32).reshape(-1), (c,)).copy()
        v = self.values[x.vid]
        vid = self.new_value(c, v.kind, v.dtype)
        self.emit(OP_ELTWISE, [x.vid], vid, i, p=dict(act=ACT_NONE), w=dict(scale=scale, shift=shift))
        self.sym[out_name] = _Sym(x.kind, vid, x.tag)

    def _op_batch_norm(self, i, op: Op):
        x = self.get(op.inp("X"))
        g, b = self.P[op.inp("Scale")], self.P[op.inp("Bias")]
        m, v = self.P[op.inp("Mean")], self.P[op.inp("Variance")]
        inv = (g.astype(np.float64) / np.sqrt(v.astype(np.float64) + op.attrs["epsilon"]))
        self._affine(i, x, inv.astype(np.float32), (b - m * inv).astype(np.float32), op.out("Y"))

    def _elementwise(self, i, op: Op, kind: str):
        xs, ys = self.get(op.inp("X")), self.get(op.inp("Y"))
        axis = op.attrs.get("axis", -1)
        out = op.out("Out")
        # attention q-scale etc. never come through here (they use `scale`)
        if xs.kind == "const" and ys.kind in ("img", "vec"):
            xs, ys = ys, xs  # commutative for add/mul
            swapped = True
        else:
            swapped = False
        if xs.kind in ("img", "vec") and ys.kind == "const":
            c = self.values[xs.vid].channels
            y = np.asarray(ys.const, np.float32)
            if y.size == 1:
                arr = np.full((c,), float(y.reshape(-1)[0]), np.float32)
            elif y.ndim == 1 and y.size == c:
                # rank-1 per-channel: axis=1 for NCHW, axis=2/-1 for [B,T,C]
                if xs.kind == "img" and xs.tag == "nchw" and axis not in (1,):
                    raise PlanError(f"per-channel operand on axis {axis} of NCHW")
                if xs.tag == "btc" and axis not in (2, -1):
                    raise PlanError(f"per-channel operand on axis {axis} of BTC")
                arr = y
            else:
                raise PlanError(f"elementwise const shape {y.shape} vs C={c}")
            if kind == "add":
                self._affine(i, xs, np.ones(c, np.float32), arr, out)
            elif kind == "mul":
                self._affine(i, xs, arr, np.zeros(c, np.float32), out)
            else:
                raise PlanError(kind)
            return
        if xs.kind == "img" and ys.kind == "vec" and kind == "mul":
            c = self.values[xs.vid].channels
            vid = self.new_value(c)
            self.emit(OP_CHSCALE, [xs.vid, ys.vid], vid, i, p=dict(residual=0))
            self.sym[out] = _Sym("img", vid, xs.tag)
            return
        if xs.kind == "img" and ys.kind == "img" and kind == "add":
            cx, cy = self.values[xs.vid].channels, self.values[ys.vid].channels
            if cx != cy:
                raise PlanError("add of different channel counts")
            if xs.tag != ys.tag:
                raise PlanError(f"add of different layouts {xs.tag} {ys.tag}")
            vid = self.new_value(cx)
            self.emit(OP_ADD, [xs.vid, ys.vid], vid, i, p=dict(act=ACT_NONE))
            self.sym[out] = _Sym("img", vid, xs.tag)
            return
        if xs.kind == "img" and ys.kind == "img" and kind == "mul":
            raise PlanError("img*img multiply")
        raise PlanError(f"elementwise_{kind} on {xs.kind},{ys.kind}")

    def _op_elementwise_add(self, i, op):
        self._elementwise(i, op, "add")

    def _op_elementwise_mul(self, i, op):
        self._elementwise(i, op, "mul")

    def _act(self, i, op: Op, act: int):
        x = self.get(op.inp("X"))
        if x.kind not in ("img", "vec"):
            raise PlanError(f"activation on {x.kind}")
        v = self.values[x.vid]
        vid = self.new_value(v.channels, v.kind, v.dtype)
        p = dict(act=act)
        if act == ACT_HSIGMOID:
            p.update(hs_slope=float(op.attrs["slope"]), hs_offset=float(op.attrs["offset"]))
        if act == ACT_HSWISH:
            a = op.attrs
            if (a["offset"], a["threshold"], a["scale"]) != (3.0, 6.0, 6.0):
                raise PlanError("non-standard hard_swish")
        if act == ACT_SWISH and op.attrs.get("beta", 1.0) != 1.0:
            raise PlanError("swish beta != 1")
        c = v.channels
        self.emit(OP_ELTWISE, [x.vid], vid, i, p=p, w=dict(scale=np.ones(c, np.float32), shift=np.zeros(c, np.float32)))
        self.nodes[-1].p["pure_act"] = 1
        self.sym[op.out("Out")] = _Sym(x.kind, vid, x.tag)

    def _op_relu(self, i, op):
        self._act(i, op, ACT_RELU)

    def _op_relu6(self, i, op):
        self._act(i, op, ACT_RELU6)

    def _op_hard_swish(self, i, op):
        self._act(i, op, ACT_HSWISH)

    def _op_hard_sigmoid(self, i, op):
        self._act(i, op, ACT_HSIGMOID)

    def _op_swish(self, i, op):
        self._act(i, op, ACT_SWISH)

    def _op_sigmoid(self, i, op):
        self._act(i, op, ACT_SIGMOID)

    def _op_pool2d(self, i, op: Op):
        a = op.attrs
        x = self.get(op.inp("X"))
        c = self.values[x.vid].channels
        ksize = list(a["ksize"])
        if (a.get("adaptive", False) and ksize == [1, 1]) or a.get("global_pooling", False):
            if a["pooling_type"] != "avg":
                raise PlanError("global max pool")
            vid = self.new_value(c, KIND_VEC, DT_F32)
            self.emit(OP_GPOOL, [x.vid], vid, i)
            self.sym[op.out("Out")] = _Sym("vec", vid)
            return
        if a.get("adaptive", False):
            raise PlanError("adaptive pool")
        pads = list(a["paddings"])
        if len(pads) == 4:
            pads = [pads[0], pads[2]]
        vid = self.new_value(c)
        self.emit(OP_POOL, [x.vid], vid, i,
                  p=dict(kh=ksize[0], kw=ksize[1], sh=a["strides"][0], sw=a["strides"][1], ph=pads[0], pw=pads[1],
                         is_max=int(a["pooling_type"] == "max"), ceil=int(bool(a.get("ceil_mode", False))),
                         exclusive=int(bool(a.get("exclusive", True)))))
        self.sym[op.out("Out")] = _Sym("img", vid)

    def _op_nearest_interp_v2(self, i, op: Op):
        a = op.attrs
        x = self.get(op.inp("X"))
        scale = list(a.get("scale", []))
        if not scale or scale[0] != scale[-1] or scale[0] != int(scale[0]) or a.get("align_corners", False):
            raise PlanError("nearest_interp: only integer uniform scale")
        c = self.values[x.vid].channels
        vid = self.new_value(c)
        self.emit(OP_UPSAMPLE, [x.vid], vid, i, p=dict(scale=int(scale[0])))
        self.sym[op.out("Out")] = _Sym("img", vid)

    def _op_concat(self, i, op: Op):
        if op.attrs["axis"] != 1:
            raise PlanError("concat axis != 1")
        ins = [self.get(n) for n in op.inputs["X"]]
        if any(s.kind != "img" for s in ins):
            raise PlanError("concat of non-image")
        ctot = sum(self.values[s.vid].channels for s in ins)
        vid = self.new_value(ctot)
        coff = 0
        for s in ins:
            c = self.values[s.vid].channels
            self.emit(OP_COPY, [s.vid], vid, i, p=dict(coff=coff, c=c))
            coff += c
        self.sym[op.out("Out")] = _Sym("img", vid)

    def _op_layer_norm(self, i, op: Op):
        x = self.get(op.inp("X"))
        if x.kind != "img" or x.tag != "btc" or op.attrs["begin_norm_axis"] != 2:
            raise PlanError("layer_norm: expects [B,T,C] over C")
        c = self.values[x.vid].channels
        vid = self.new_value(c)
        self.emit(OP_LAYERNORM, [x.vid], vid, i, p=dict(eps=float(op.attrs["epsilon"])),
                  w=dict(gamma=self.P[op.inp("Scale")].astype(np.float32), beta=self.P[op.inp("Bias")].astype(np.float32)))
        self.sym[op.out("Y")] = _Sym("img", vid, "btc")

    def _matmul(self, i, op: Op):
        a = op.attrs
        xs, ys = self.get(op.inp("X")), self.get(op.inp("Y"))
        tx = a.get("trans_x", a.get("transpose_X", False))
        ty = a.get("trans_y", a.get("transpose_Y", False))
        out = op.out("Out")
        if xs.kind == "img" and ys.kind == "const":
            if xs.tag != "btc" or tx:
                raise PlanError("matmul: lhs must be [B,T,C]")
            w = np.asarray(ys.const, np.float32)
            if ty:
                w = w.T
            if op.type == "matmul" and a.get("alpha", 1.0) != 1.0:
                w = w * a["alpha"]
            k, n = w.shape
            if k != self.values[xs.vid].channels:
                raise PlanError("matmul K mismatch")
            vid = self.new_value(n)
            self.emit(OP_CONV, [xs.vid], vid, i, p=dict(kh=1, kw=1, sh=1, sw=1, ph=0, pw=0, cin=k, cout=n),
                      w=dict(weight=np.ascontiguousarray(w.T).reshape(n, 1, 1, k)))
            self.sym[out] = _Sym("img", vid, "btc")
            return
        if xs.kind == "q" and ys.kind == "kT":
            if xs.extra["qkv"] != ys.extra["qkv"]:
                raise PlanError("attention q/k from different projections")
            self.sym[out] = _Sym("scores", extra=dict(xs.extra))
            return
        if xs.kind == "probs" and ys.kind == "v":
            if xs.extra["qkv"] != ys.extra["qkv"]:
                raise PlanError("attention p/v from different projections")
            self.sym[out] = _Sym("ctx", extra=dict(xs.extra))
            return
        raise PlanError(f"matmul on {xs.kind},{ys.kind}")

    def _op_matmul_v2(self, i, op):
        self._matmul(i, op)

    def _op_matmul(self, i, op):
        self._matmul(i, op)

    def _op_mul(self, i, op):
        self._matmul(i, op)

    def _op_scale(self, i, op: Op):
        a = op.attrs
        x = self.get(op.inp("X"))
        s, b = float(a.get("scale", 1.0)), float(a.get("bias", 0.0))
        out = op.out("Out")
        if x.kind == "q" or (x.kind == "qkv_slice" and x.extra["which"] == 0):
            if b != 0.0:
                raise PlanError("attention scale with bias")
            e = dict(x.extra)
            e["qscale"] = e.get("qscale", 1.0) * s
            self.sym[out] = _Sym("q", extra=e)
            return
        if x.kind == "scores":
            if b != 0.0:
                raise PlanError("attention scale with bias")
            e = dict(x.extra)
            e["qscale"] = e.get("qscale", 1.0) * s
            self.sym[out] = _Sym("scores", extra=e)
            return
        if x.kind in ("img", "vec"):
            if s == 1.0 and b == 0.0:
                self.sym[out] = x
                return
            c = self.values[x.vid].channels
            if not a.get("bias_after_scale", True):
                b = b * s
            self._affine(i, x, np.full(c, s, np.float32), np.full(c, b, np.float32), out)
            return
        if x.kind == "const":
            self.sym[out] = _Sym("const", const=np.asarray(x.const) * s + b)
            return
        raise PlanError(f"scale on {x.kind}")

    def _op_softmax(self, i, op: Op):
        x = self.get(op.inp("X"))
        out = op.out("Out")
        if x.kind == "scores":
            self.sym[out] = _Sym("probs", extra=dict(x.extra))
            return
        if x.kind == "img" and x.tag == "btc" and op.attrs.get("axis", -1) in (-1, 2):
            c = self.values[x.vid].channels
            vid = self.new_value(c, KIND_IMG, DT_F32)
            self.emit(OP_SOFTMAX, [x.vid], vid, i)
            self.sym[out] = _Sym("img", vid, "btc")
            return
        raise PlanError(f"softmax on {x.kind}/{x.tag}")

    def _op_dropout(self, i, op: Op):
        self.sym[op.out("Out")] = self.get(op.inp("X"))

    def _op_assign(self, i, op: Op):
        self.sym[op.out("Out")] = self.get(op.inp("X"))

    def _op_shape(self, i, op: Op):
        x = self.get(op.inp("Input"))
        self.sym[op.out("Out")] = _Sym("shapeof", vid=x.vid, tag=x.tag)

    def _op_fill_constant(self, i, op: Op):
        a = op.attrs
        val = a.get("str_value", "") or a["value"]
        self.sym[op.out("Out")] = _Sym("const", const=np.full(list(a["shape"]) or [1], float(val)))

    def _op_fill_constant_batch_size_like(self, i, op: Op):
        self.sym[op.out("Out")] = _Sym("zeros_state")

    def _op_slice(self, i, op: Op):
        a = op.attrs
        x = self.get(op.inp("Input"))
        out = op.out("Out")
        if x.kind == "shapeof":
            self.sym[out] = _Sym("shapedim", extra=dict(axis=a["starts"][0]))
            return
        if x.kind == "qkv_t":
            if a["axes"] != [0] or a["ends"][0] - a["starts"][0] != 1 or a.get("decrease_axis", []) != [0]:
                raise PlanError("unexpected qkv slice")
            e = dict(x.extra)
            e["which"] = a["starts"][0]
            kind = {0: "q", 1: "k", 2: "v"}[e["which"]]
            self.sym[out] = _Sym(kind, extra=e)
            return
        raise PlanError(f"slice on {x.kind}")

    def _op_flatten_contiguous_range(self, i, op: Op):
        x = self.get(op.inp("X"))
        a = op.attrs
        if x.kind == "img" and x.tag == "nchw" and (a["start_axis"], a["stop_axis"]) == (2, 3):
            self.sym[op.out("Out")] = _Sym("img", x.vid, "bcw", extra=dict(need_h1=True))
            self._need_h1(x.vid)
            return
        raise PlanError("flatten")

    def _need_h1(self, vid):
        if not hasattr(self, "h1_values"):
            self.h1_values = set()
        self.h1_values.add(vid)

    def _op_squeeze2(self, i, op: Op):
        x = self.get(op.inp("X"))
        if x.kind == "img" and x.tag == "nchw" and list(op.attrs["axes"]) == [2]:
            self._need_h1(x.vid)
            self.sym[op.out("Out")] = _Sym("img", x.vid, "bcw")
            return
        raise PlanError("squeeze2")

    def _op_transpose2(self, i, op: Op):
        x = self.get(op.inp("X"))
        axis = list(op.attrs["axis"])
        out = op.out("Out")
        if x.kind == "img":
            table = {("bcw", (0, 2, 1)): "btc", ("btc", (0, 2, 1)): "bcw", ("b1tc", (0, 3, 1, 2)): "nchw",
                     ("bcw", (2, 0, 1)): "tbc", ("tbc", (1, 0, 2)): "btc", ("btc", (1, 0, 2)): "tbc", ("nchw", (0, 2, 3, 1)): "b1tc"}
            key = (x.tag, tuple(axis))
            if key not in table:
                raise PlanError(f"transpose {key}")
            self.sym[out] = _Sym("img", x.vid, table[key])
            return
        if x.kind == "qkv" and axis == [2, 0, 3, 1, 4]:
            self.sym[out] = _Sym("qkv_t", extra=dict(x.extra))
            return
        if x.kind == "k" and axis == [0, 1, 3, 2]:
            self.sym[out] = _Sym("kT", extra=dict(x.extra))
            return
        if x.kind == "ctx" and axis == [0, 2, 1, 3]:
            self.sym[out] = _Sym("ctx_t", extra=dict(x.extra))
            return
        raise PlanError(f"transpose2 on {x.kind} {axis}")

    def _op_reshape2(self, i, op: Op):
        x = self.get(op.inp("X"))
        out = op.out("Out")
        shape = list(op.attrs.get("shape", []))
        if x.kind == "img" and x.tag == "btc":
            c = self.values[x.vid].channels
            if len(shape) == 5 and shape[:2] == [0, -1] and shape[2] == 3 and shape[2] * shape[3] * shape[4] == c:
                self.sym[out] = _Sym("qkv", extra=dict(qkv=x.vid, heads=shape[3], dim=shape[4]))
                return
            if len(shape) == 4 and shape[1] == 1 and shape[3] == c:  # [B,1,T,C]
                self.sym[out] = _Sym("img", x.vid, "b1tc")
                return
            if len(shape) == 3 and shape[2] == c:
                self.sym[out] = x
                return
        if x.kind == "ctx_t":
            e = x.extra
            if len(shape) == 3 and shape[2] == e["heads"] * e["dim"]:
                c = e["heads"] * e["dim"]
                vid = self.new_value(c)
                self.emit(OP_ATTN, [e["qkv"]], vid, i, p=dict(heads=e["heads"], dim=e["dim"], qscale=float(e.get("qscale", 1.0))))
                self.sym[out] = _Sym("img", vid, "btc")
                return
        raise PlanError(f"reshape2 on {x.kind}/{x.tag} to {shape}")

    def _op_rnn(self, i, op: Op):
        a = op.attrs
        x = self.get(op.inp("Input"))
        if x.kind != "img" or x.tag != "tbc" or a["mode"] != "LSTM":
            raise PlanError("rnn: expects time-major LSTM")
        hidden, layers, bidir = a["hidden_size"], a["num_layers"], a["is_bidirec"]
        ndir = 2 if bidir else 1
        wl = [self.P[n] for n in op.inputs["WeightList"]]
        nw = layers * ndir
        # One layer = the input GEMM of both directions as ONE 1x1 convolution (cin -> ndir*4*hidden gate pre-activations,
        # b_ih + b_hh folded into its bias; runs on the tensor-core conv path) followed by the recurrent LSTM step, which
        # only carries W_hh [ndir][4*hidden][hidden] (gate order i,f,g,o) and walks the time axis.
        cur = x.vid
        for l in range(layers):
            cin = self.values[cur].channels
            w_ih = np.concatenate([wl[2 * (l * ndir + d)] for d in range(ndir)], 0).astype(np.float32)        # [ndir*4H, cin]
            w_hh = np.stack([wl[2 * (l * ndir + d) + 1] for d in range(ndir)], 0).astype(np.float32)          # [ndir, 4H, H]
            b = np.concatenate([wl[2 * nw + 2 * (l * ndir + d)] + wl[2 * nw + 2 * (l * ndir + d) + 1] for d in range(ndir)], 0)
            if w_ih.shape != (ndir * 4 * hidden, cin) or w_hh.shape != (ndir, 4 * hidden, hidden):
                raise PlanError("rnn: unexpected weight shapes")
            gates = self.new_value(ndir * 4 * hidden)
            self.emit(OP_CONV, [cur], gates, i, p=dict(kh=1, kw=1, sh=1, sw=1, ph=0, pw=0, cin=cin, cout=ndir * 4 * hidden),
                      w=dict(weight=w_ih.reshape(ndir * 4 * hidden, 1, 1, cin), bias=b.astype(np.float32)))
            vid = self.new_value(hidden * ndir)
            self.emit(OP_LSTM, [gates], vid, i, p=dict(hidden=hidden, ndir=ndir, cin=ndir * 4 * hidden, cout=hidden * ndir),
                      w=dict(weight=w_hh))
            cur = vid
        self.sym[op.out("Out")] = _Sym("img", vid, "tbc")


# --------------------------------------------------------------------------- #
# fusion
# --------------------------------------------------------------------------- #

_EPI_OPS = (OP_CONV, OP_DWCONV, OP_DECONV2, OP_VECLIN, OP_STEM)


def _consumers(nodes: List[Step]) -> Dict[int, List[int]]:
    cons: Dict[int, List[int]] = {}
    for idx, s in enumerate(nodes):
        for v in s.ins:
            cons.setdefault(v, []).append(idx)
    return cons


def _fuse(values: List[Value], nodes: List[Step], keep: set) -> List[Step]:
    """Fold ELTWISE/ADD chains into conv epilogues; UPSAMPLE+ADD; CHSCALE+ADD (RSE); ADD+act."""
    cons = _consumers(nodes)
    dead = [False] * len(nodes)
    defined_at = {s.out: idx for idx, s in enumerate(nodes) if s.op != OP_COPY}

    def sole_consumer(vid: int) -> Optional[int]:
        c = cons.get(vid, [])
        if len(c) == 1 and vid not in keep:
            return c[0]
        return None

    for idx, s in enumerate(nodes):
        if dead[idx]:
            continue
        if s.op in _EPI_OPS:
            cout = values[s.out].channels
            s1 = np.ones(cout, np.float64)
            b1 = s.w["bias"].astype(np.float64) if "bias" in s.w else np.zeros(cout, np.float64)   # e.g. the LSTM input GEMM
            s2 = np.ones(cout, np.float64)
            b2 = np.zeros(cout, np.float64)
            act1, act2, res = ACT_NONE, ACT_NONE, -1
            hs = (0.0, 0.0)
            stage = 0  # 0: pre-act, 1: post-act1, 2: post-residual
            cur = s.out
            while True:
                nxt = sole_consumer(cur)
                if nxt is None or dead[nxt]:
                    break
                n = nodes[nxt]
                if n.op == OP_ELTWISE and n.ins[0] == cur:
                    sc, sf, act = n.w["scale"].astype(np.float64), n.w["shift"].astype(np.float64), n.p["act"]
                    if stage == 0:
                        s1, b1 = s1 * sc, b1 * sc + sf
                        if act != ACT_NONE:
                            act1, hs, stage = act, (n.p.get("hs_slope", 0.0), n.p.get("hs_offset", 0.0)), 1
                    elif stage == 1:
                        if act != ACT_NONE:
                            break
                        s2, b2 = s2 * sc, b2 * sc + sf
                    else:
                        if act2 != ACT_NONE or not n.p.get("pure_act") or act in (ACT_HSIGMOID,):
                            break
                        act2 = act
                elif n.op == OP_ADD and res < 0 and stage < 2 and s.op != OP_VECLIN:
                    other = n.ins[1] if n.ins[0] == cur else n.ins[0]
                    # the residual operand must already exist when this conv runs
                    if other == cur or defined_at.get(other, -1) > idx:
                        break
                    res, stage = other, 2
                else:
                    break
                dead[nxt] = True
                s.src_ops += n.src_ops
                cur = n.out
            s.out = cur
            # fold s1 into the weights, keep bias
            w = s.w["weight"].astype(np.float64)
            if s.op in (OP_CONV, OP_STEM):
                w = w * s1.reshape(-1, 1, 1, 1)
            elif s.op == OP_DWCONV:
                w = w * s1.reshape(1, 1, -1)
            elif s.op == OP_DECONV2:
                w = w * s1.reshape(1, 1, -1, 1)
            else:
                w = w * s1.reshape(-1, 1)
            s.w["weight"] = w.astype(np.float32)
            s.w["bias"] = b1.astype(np.float32)
            s.p.update(act=act1, act2=act2, hs_slope=float(hs[0]), hs_offset=float(hs[1]),
                       has_post=int(not (np.all(s2 == 1.0) and np.all(b2 == 0.0))))
            if s.p["has_post"]:
                s.w["post_scale"] = s2.astype(np.float32)
                s.w["post_shift"] = b2.astype(np.float32)
            if res >= 0:
                s.ins = [s.ins[0], res]
            s.p["has_res"] = int(res >= 0)
        elif s.op == OP_UPSAMPLE:
            nxt = sole_consumer(s.out)
            if nxt is not None and nodes[nxt].op == OP_ADD and not dead[nxt]:
                n = nodes[nxt]
                other = n.ins[1] if n.ins[0] == s.out else n.ins[0]
                if other in defined_at and defined_at[other] < idx:
                    s.ins = [s.ins[0], other]
                    s.p["has_add"] = 1
                    s.out = n.out
                    s.src_ops += n.src_ops
                    dead[nxt] = True
            s.p.setdefault("has_add", 0)
        elif s.op == OP_CHSCALE:
            nxt = sole_consumer(s.out)
            if nxt is not None and nodes[nxt].op == OP_ADD and not dead[nxt]:
                n = nodes[nxt]
                other = n.ins[1] if n.ins[0] == s.out else n.ins[0]
                if other == s.ins[0]:  # x + x*s  (RSE)
                    s.p["residual"] = 1
                    s.out = n.out
                    s.src_ops += n.src_ops
                    dead[nxt] = True
        elif s.op == OP_ADD:
            nxt = sole_consumer(s.out)
            if nxt is not None and nodes[nxt].op == OP_ELTWISE and nodes[nxt].p.get("pure_act") and not dead[nxt] \
                    and nodes[nxt].p["act"] != ACT_HSIGMOID:
                s.p["act"] = nodes[nxt].p["act"]
                s.out = nodes[nxt].out
                s.src_ops += nodes[nxt].src_ops
                dead[nxt] = True
    return [s for idx, s in enumerate(nodes) if not dead[idx]]


def _reorder_concats(values: List[Value], nodes: List[Step], keep: set) -> None:
    """Concat buffers that are only read by dense convolutions may hold their slices in any order, as long as the
    convolutions' input channels are permuted the same way.  Slices whose width is a multiple of CPAD go first, so that
    their producers can write them in place (`_resolve_concat` needs CPAD-aligned offsets): PFHeadLocal of the server
    detector concatenates a 1-channel map BEFORE a 64-channel one (V4/ch_det, ops #318-321) — in graph order the 64-channel
    full-resolution slice starts at channel 1 and costs a copy of the largest activation of the whole network."""
    copies: Dict[int, List[Step]] = {}
    for s in nodes:
        if s.op == OP_COPY:
            copies.setdefault(s.out, []).append(s)
    cons = _consumers(nodes)
    for root, cps in copies.items():
        if root in keep or len(cps) < 2:
            continue
        users = [nodes[i] for i in cons.get(root, []) if nodes[i].op != OP_COPY or nodes[i].out != root]
        if not users or any(u.op != OP_CONV or u.ins[0] != root or root in u.ins[1:] for u in users):
            continue
        cps_sorted = sorted(cps, key=lambda c: c.p["coff"])
        if all(c.p["coff"] % CPAD == 0 for c in cps_sorted):
            continue
        order = sorted(cps_sorted, key=lambda c: 0 if c.p["c"] % CPAD == 0 else 1)       # stable
        perm: List[int] = []                                                             # new channel -> old channel
        off = 0
        for c in order:
            perm += list(range(c.p["coff"], c.p["coff"] + c.p["c"]))
            c.p["coff"] = off
            off += c.p["c"]
        if sorted(perm) != list(range(values[root].channels)):
            raise PlanError("concat slices do not tile the buffer")
        for u in users:
            u.w["weight"] = np.ascontiguousarray(u.w["weight"][..., perm])                # [cout][kh][kw][cin]


def _resolve_concat(values: List[Value], nodes: List[Step], input_vid: int) -> List[Step]:
    """Turn COPY-into-concat into aliasing when the producer can write the slice directly."""
    producers: Dict[int, int] = {}
    for idx, s in enumerate(nodes):
        if s.op != OP_COPY:
            producers[s.out] = idx
    cons = _consumers(nodes)
    out: List[Step] = []
    for idx, s in enumerate(nodes):
        if s.op == OP_COPY:
            src, dst, coff = s.ins[0], s.out, s.p["coff"]
            v = values[src]
            ok = (coff % CPAD == 0 and v.alias_of < 0 and src != input_vid and src in producers
                  and v.dtype == values[dst].dtype and v.kind == KIND_IMG
                  and (v.channels % CPAD == 0 or coff + v.channels == values[dst].channels))
            # a source that is itself consumed as a residual/input elsewhere is fine: views are readable
            if ok:
                v.alias_of, v.alias_coff = dst, coff
                continue
        out.append(s)
    return out


# --------------------------------------------------------------------------- #
# plan object + serialization
# --------------------------------------------------------------------------- #

PLAN_MAGIC = 0x50455356  # 'VSEP'
PLAN_VERSION = 3
_N_INS, _N_P, _N_F, _N_W = 4, 20, 4, 8

# integer parameter slots per op (mirrored in csrc/plan.h)
_P_SLOTS = ["kh", "kw", "sh", "sw", "ph", "pw", "cin", "cout", "act", "act2", "has_post", "has_res", "scale", "has_add",
            "residual", "is_max", "ceil", "exclusive", "heads", "dim"]
_F_SLOTS = ["hs_slope", "hs_offset", "eps", "qscale"]
_W_SLOTS = ["weight", "bias", "post_scale", "post_shift", "gamma", "beta", "scale", "shift"]
_COPY_P = {"coff": "scale", "c": "cout"}           # COPY reuses slots
_LSTM_P = {"hidden": "heads", "ndir": "scale"}


@dataclass
class Plan:
    values: List[Value]
    steps: List[Step]
    input_vid: int
    output_vids: List[int]
    name: str = ""
    norm_scale: Tuple[float, float, float] = (1.0, 1.0, 1.0)   # u8 -> float input normalisation (per BGR channel)
    norm_shift: Tuple[float, float, float] = (0.0, 0.0, 0.0)
    h1_values: Tuple[int, ...] = ()

    def summary(self) -> str:
        lines = [f"plan {self.name}: {len(self.steps)} steps, {len(self.values)} values"]
        for k, s in enumerate(self.steps):
            lines.append(f"  {k:3d} {s!r}")
        return "\n".join(lines)

    def liveness(self) -> Dict[int, Tuple[int, int]]:
        """value id -> (first def step, last use step) with alias roots merged."""
        root = lambda v: self.values[v].alias_of if self.values[v].alias_of >= 0 else v
        live: Dict[int, List[int]] = {}
        for k, s in enumerate(self.steps):
            for v in list(s.ins) + [s.out]:
                r = root(v)
                if r not in live:
                    live[r] = [k, k]
                live[r][1] = k
        for v in self.output_vids + [self.input_vid]:
            r = root(v)
            if r in live:
                live[r][1] = len(self.steps)
            else:
                live[r] = [0, len(self.steps)]
        live[root(self.input_vid)][0] = -1
        return {k: (a, b) for k, (a, b) in live.items()}

    def serialize(self) -> bytes:
        wchunks: List[np.ndarray] = []
        woff = 0

        def add_w(arr: np.ndarray) -> int:
            nonlocal woff
            arr = np.ascontiguousarray(arr, dtype=np.float32).reshape(-1)
            pad = (-arr.size) % 4
            if pad:
                arr = np.concatenate([arr, np.zeros(pad, np.float32)])
            off = woff
            wchunks.append(arr)
            woff += arr.size
            return off

        live = self.liveness()
        vrec = b""
        for v in self.values:
            root = v.alias_of if v.alias_of >= 0 else v.vid
            a, b = live.get(root, (-2, -2))   # -2: value eliminated by fusion, never materialised
            vrec += struct.pack("<8i", v.channels, v.cstride, v.kind, v.dtype, v.alias_of, v.alias_coff, a, b)
        srec = b""
        for s in self.steps:
            ins = (list(s.ins) + [-1] * _N_INS)[:_N_INS]
            p = [0] * _N_P
            f = [0.0] * _N_F
            wslots = [-1] * _N_W
            wsizes = [0] * _N_W
            remap = _COPY_P if s.op == OP_COPY else (_LSTM_P if s.op == OP_LSTM else {})
            for k, val in s.p.items():
                k = remap.get(k, k)
                if k in _P_SLOTS:
                    p[_P_SLOTS.index(k)] = int(val)
                elif k in _F_SLOTS:
                    f[_F_SLOTS.index(k)] = float(val)
                elif k in ("pure_act",):
                    pass
                else:
                    raise PlanError(f"unknown step param {k}")
            for k, arr in s.w.items():
                j = _W_SLOTS.index(k)
                wslots[j], wsizes[j] = add_w(arr), int(np.asarray(arr).size)
            srec += struct.pack(f"<i{_N_INS}ii{_N_P}i{_N_F}f{_N_W}q{_N_W}q", s.op, *ins, s.out, *p, *f, *wslots, *wsizes)
        weights = np.concatenate(wchunks) if wchunks else np.zeros(0, np.float32)
        name_b = self.name.encode()[:63].ljust(64, b"\0")
        outs = (list(self.output_vids) + [-1] * 4)[:4]
        h1 = (list(self.h1_values) + [-1] * 8)[:8]
        header = struct.pack("<6I", PLAN_MAGIC, PLAN_VERSION, len(self.values), len(self.steps), 0, 0)
        header += struct.pack("<q", weights.size)
        header += struct.pack("<i4i8i", self.input_vid, *outs, *h1)
        header += struct.pack("<6f", *self.norm_scale, *self.norm_shift)
        header += name_b
        return header + vrec + srec + weights.tobytes()


def compile_model(model: Model, name: str = "", norm_scale=(1.0, 1.0, 1.0), norm_shift=(0.0, 0.0, 0.0),
                  fetch_cols: Optional[List[int]] = None) -> Plan:
    low = _Lowerer(model).run()
    fetch = getattr(low, "fetch_vids", {})
    cols = sorted(fetch) if fetch_cols is None else fetch_cols
    out_vids = [fetch[c] for c in cols]
    # dead code elimination w.r.t. the requested fetch columns
    nodes = low.nodes
    needed = set(out_vids)
    keep_nodes = [False] * len(nodes)
    for idx in range(len(nodes) - 1, -1, -1):
        if nodes[idx].out in needed:
            keep_nodes[idx] = True
            needed.update(nodes[idx].ins)
    nodes = [n for n, k in zip(nodes, keep_nodes) if k]
    nodes = _fuse(low.values, nodes, keep=set(out_vids))
    _reorder_concats(low.values, nodes, set(out_vids))
    nodes = _resolve_concat(low.values, nodes, low.input_vid)
    # fetched values leave the engine as dense float32 [pixels][channels]
    cons = _consumers(nodes)
    producer = {s.out: s for s in nodes if s.op != OP_COPY}
    for k, v in enumerate(list(out_vids)):
        val = low.values[v]
        if val.dtype == DT_F32:
            continue
        direct = (v in producer and producer[v].op in (OP_CONV, OP_DECONV2, OP_ELTWISE, OP_ADD) and not cons.get(v)
                  and val.alias_of < 0 and not any(x.alias_of == v for x in low.values))
        if direct:
            val.dtype = DT_F32
        else:
            c = val.channels
            nv = Value(len(low.values), c, KIND_IMG, DT_F32)
            low.values.append(nv)
            nodes.append(Step(OP_ELTWISE, [v], nv.vid, dict(act=ACT_NONE),
                              dict(scale=np.ones(c, np.float32), shift=np.zeros(c, np.float32))))
            out_vids[k] = nv.vid
    return Plan(values=low.values, steps=nodes, input_vid=low.input_vid, output_vids=out_vids, name=name,
                norm_scale=tuple(norm_scale), norm_shift=tuple(norm_shift),
                h1_values=tuple(sorted(getattr(low, "h1_values", set()))))


# det: (x/255 - mean)/std per BGR channel; rec: (x/255 - 0.5)/0.5     
DET_NORM = (tuple(1.0 / (255.0 * s) for s in (0.229, 0.224, 0.225)),
            tuple(-m / s for m, s in zip((0.485, 0.456, 0.406), (0.229, 0.224, 0.225))))
REC_NORM = ((1.0 / 127.5,) * 3, (-1.0,) * 3)


def blob_step_count(blob: bytes) -> int:
    """Number of steps of a packed plan (header field n_steps, csrc/plan.h::PlanHeader)."""
    import struct
    magic, version, n_values, n_steps = struct.unpack_from("<4I", blob, 0)
    if magic != PLAN_MAGIC:
        raise ValueError("not a packed plan")
    return int(n_steps)


def deserialize(blob: bytes) -> Plan:
    """Inverse of ``Plan.serialize`` (tests rebuild the step list from a shipped .vsep file)."""
    hdr_fmt = "<6Iqi4i8i6f64s"
    hsz = struct.calcsize(hdr_fmt)
    f = struct.unpack_from(hdr_fmt, blob, 0)
    magic, version, n_values, n_steps = f[0], f[1], f[2], f[3]
    if magic != PLAN_MAGIC or version != PLAN_VERSION:
        raise PlanError("bad plan blob")
    n_weights, input_vid = f[6], f[7]
    outs = [v for v in f[8:12] if v >= 0]
    h1 = [v for v in f[12:20] if v >= 0]
    norm = f[20:26]
    name = f[26].split(b"\0")[0].decode()
    pos = hsz
    values: List[Value] = []
    for vid in range(n_values):
        ch, cs, kind, dtype, alias_of, alias_coff, a, b = struct.unpack_from("<8i", blob, pos)
        pos += 32
        values.append(Value(vid, ch, kind, dtype, alias_of, alias_coff))
    step_fmt = f"<i{_N_INS}ii{_N_P}i{_N_F}f{_N_W}q{_N_W}q"
    ssz = struct.calcsize(step_fmt)
    wbase = pos + ssz * n_steps
    weights = np.frombuffer(blob, dtype=np.float32, count=n_weights, offset=wbase)
    steps: List[Step] = []
    for _ in range(n_steps):
        r = struct.unpack_from(step_fmt, blob, pos)
        pos += ssz
        op = r[0]
        ins = [v for v in r[1:1 + _N_INS] if v >= 0]
        out = r[1 + _N_INS]
        pv = r[2 + _N_INS:2 + _N_INS + _N_P]
        fv = r[2 + _N_INS + _N_P:2 + _N_INS + _N_P + _N_F]
        wo = r[2 + _N_INS + _N_P + _N_F:2 + _N_INS + _N_P + _N_F + _N_W]
        wn = r[2 + _N_INS + _N_P + _N_F + _N_W:]
        p: Dict[str, Any] = {k: pv[i] for i, k in enumerate(_P_SLOTS)}
        p.update({k: fv[i] for i, k in enumerate(_F_SLOTS)})
        if op == OP_COPY:
            p["coff"], p["c"] = p["scale"], p["cout"]
        if op == OP_LSTM:
            p["hidden"], p["ndir"] = p["heads"], p["scale"]
        w: Dict[str, np.ndarray] = {}
        for j, k in enumerate(_W_SLOTS):
            if wo[j] >= 0:
                w[k] = weights[wo[j]:wo[j] + wn[j]]
        cin, cout, kh, kw = p["cin"], p["cout"], p["kh"], p["kw"]
        if op in (OP_CONV, OP_STEM):
            w["weight"] = w["weight"].reshape(cout, kh, kw, cin)
        elif op == OP_DWCONV:
            w["weight"] = w["weight"].reshape(kh, kw, cin)
        elif op == OP_DECONV2:
            w["weight"] = w["weight"].reshape(2, 2, cout, cin)
        elif op == OP_VECLIN:
            w["weight"] = w["weight"].reshape(cout, cin)
        elif op == OP_LSTM:
            w["weight"] = w["weight"].reshape(p["ndir"], 4 * p["hidden"], p["hidden"])
        steps.append(Step(op, ins, out, p, w))
    return Plan(values=values, steps=steps, input_vid=input_vid, output_vids=outs, name=name,
                norm_scale=tuple(norm[:3]), norm_shift=tuple(norm[3:]), h1_values=tuple(h1))
