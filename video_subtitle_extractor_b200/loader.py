"""Reader for the reference's shipped Paddle inference models.

The reference loads ``backend/models/<ver>/<name>/inference.pdmodel`` +
``inference.pdiparams`` through ``paddle.inference`` (reference:
backend/tools/paddle_model_config.py:8-106 selects the directory;
backend/tools/ocr.py:91-113 hands it to PaddleOCR).  Paddle is not part of this
engine, so the two container formats are parsed directly:

* ``inference.pdmodel``  = a serialized ``ProgramDesc`` protobuf (proto2),
  decoded with the ~100-line wire-format reader below.
* ``inference.pdiparams`` = one record per persistable variable, in ascending
  name order: ``u32 0 | u64 0 | u32 0 | i32 desc_len | TensorDesc | raw fp32``.
  Split parameter files (``inference_N.pdiparams`` + ``fs_manifest.csv``) are
  concatenated in memory – the reference merges them on disk with ``fsplit``
  (paddle_model_config.py:100-106), which a read-only model tree forbids.
"""
from __future__ import annotations

import csv
import os
import struct
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional, Tuple

import numpy as np

# --------------------------------------------------------------------------- #
# protobuf wire format
# --------------------------------------------------------------------------- #


def _varint(buf: bytes, pos: int) -> Tuple[int, int]:
    result = 0
    shift = 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7


def _fields(buf: bytes):
    """Yield (field_number, wire_type, value) over one message."""
    pos = 0
    n = len(buf)
    while pos < n:
        key, pos = _varint(buf, pos)
        fno, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _varint(buf, pos)
        elif wt == 1:
            v = buf[pos:pos + 8]
            pos += 8
        elif wt == 2:
            ln, pos = _varint(buf, pos)
            v = buf[pos:pos + ln]
            pos += ln
        elif wt == 5:
            v = buf[pos:pos + 4]
            pos += 4
        else:
            raise ValueError(f"unsupported wire type {wt}")
        yield fno, wt, v


def _s64(v: int) -> int:
    return v - (1 << 64) if v >= (1 << 63) else v


def _s32(v: int) -> int:
    v &= 0xFFFFFFFFFFFFFFFF
    v = _s64(v)
    return v


def _packed_varints(buf: bytes) -> List[int]:
    out = []
    pos = 0
    while pos < len(buf):
        v, pos = _varint(buf, pos)
        out.append(_s64(v))
    return out


# --------------------------------------------------------------------------- #
# ProgramDesc model
# --------------------------------------------------------------------------- #

_DTYPES = {0: np.bool_, 1: np.int16, 2: np.int32, 3: np.int64, 4: np.float16, 5: np.float32, 6: np.float64}


@dataclass
class Var:
    name: str
    persistable: bool = False
    dtype: Optional[int] = None
    dims: List[int] = field(default_factory=list)


@dataclass
class Op:
    type: str
    inputs: Dict[str, List[str]] = field(default_factory=dict)
    outputs: Dict[str, List[str]] = field(default_factory=dict)
    attrs: Dict[str, Any] = field(default_factory=dict)

    def inp(self, key: str, idx: int = 0) -> str:
        return self.inputs[key][idx]

    def out(self, key: str, idx: int = 0) -> str:
        return self.outputs[key][idx]


@dataclass
class Program:
    vars: Dict[str, Var]
    ops: List[Op]
    version: int = 0

    @property
    def feed_names(self) -> List[str]:
        feeds = [(op.attrs.get("col", 0), op.out("Out")) for op in self.ops if op.type == "feed"]
        return [n for _, n in sorted(feeds)]

    @property
    def fetch_names(self) -> List[str]:
        fetches = [(op.attrs.get("col", 0), op.inp("X")) for op in self.ops if op.type == "fetch"]
        return [n for _, n in sorted(fetches)]


def _parse_tensor_desc(buf: bytes) -> Tuple[int, List[int]]:
    dtype = None
    dims: List[int] = []
    for fno, wt, v in _fields(buf):
        if fno == 1:
            dtype = v
        elif fno == 2:
            if wt == 2:
                dims.extend(_packed_varints(v))
            else:
                dims.append(_s64(v))
    return dtype, dims


def _parse_var(buf: bytes) -> Var:
    var = Var(name="")
    for fno, wt, v in _fields(buf):
        if fno == 1:
            var.name = v.decode()
        elif fno == 3:
            var.persistable = bool(v)
        elif fno == 2:  # VarType
            for f2, _, v2 in _fields(v):
                if f2 == 3:  # LoDTensorDesc
                    for f3, _, v3 in _fields(v2):
                        if f3 == 1:
                            var.dtype, var.dims = _parse_tensor_desc(v3)
    return var


def _parse_attr(buf: bytes) -> Tuple[str, Any]:
    name = ""
    atype = None
    scalars: Dict[int, Any] = {}
    ints: List[int] = []
    floats: List[float] = []
    strings: List[str] = []
    bools: List[bool] = []
    longs: List[int] = []
    for fno, wt, v in _fields(buf):
        if fno == 1:
            name = v.decode()
        elif fno == 2:
            atype = v
        elif fno == 3:
            scalars[3] = _s32(v)
        elif fno == 4:
            scalars[4] = struct.unpack("<f", v)[0]
        elif fno == 5:
            scalars[5] = v.decode(errors="replace")
        elif fno == 6:
            ints.extend(_packed_varints(v) if wt == 2 else [_s32(v)])
        elif fno == 7:
            if wt == 2:
                floats.extend(struct.unpack(f"<{len(v) // 4}f", v))
            else:
                floats.append(struct.unpack("<f", v)[0])
        elif fno == 8:
            strings.append(v.decode(errors="replace"))
        elif fno == 10:
            scalars[10] = bool(v)
        elif fno == 11:
            bools.extend([bool(x) for x in _packed_varints(v)] if wt == 2 else [bool(v)])
        elif fno == 12:
            scalars[12] = v  # block idx
        elif fno == 13:
            scalars[13] = _s64(v)
        elif fno == 15:
            longs.extend(_packed_varints(v) if wt == 2 else [_s64(v)])
    value: Any
    if atype == 0:
        value = scalars.get(3, 0)
    elif atype == 1:
        value = scalars.get(4, 0.0)
    elif atype == 2:
        value = scalars.get(5, "")
    elif atype == 3:
        value = ints
    elif atype == 4:
        value = floats
    elif atype == 5:
        value = strings
    elif atype == 6:
        value = scalars.get(10, False)
    elif atype == 7:
        value = bools
    elif atype == 8:
        value = scalars.get(12, 0)
    elif atype == 9:
        value = scalars.get(13, 0)
    elif atype == 11:
        value = longs
    else:
        value = None
    return name, value


def _parse_op_var(buf: bytes) -> Tuple[str, List[str]]:
    param = ""
    args: List[str] = []
    for fno, _, v in _fields(buf):
        if fno == 1:
            param = v.decode()
        elif fno == 2:
            args.append(v.decode())
    return param, args


def _parse_op(buf: bytes) -> Op:
    op = Op(type="")
    for fno, _, v in _fields(buf):
        if fno == 3:
            op.type = v.decode()
        elif fno == 1:
            k, a = _parse_op_var(v)
            op.inputs[k] = a
        elif fno == 2:
            k, a = _parse_op_var(v)
            op.outputs[k] = a
        elif fno == 4:
            k, a = _parse_attr(v)
            op.attrs[k] = a
    return op


def parse_program(buf: bytes) -> Program:
    vars_: Dict[str, Var] = {}
    ops: List[Op] = []
    version = 0
    n_blocks = 0
    for fno, _, v in _fields(buf):
        if fno == 1:  # BlockDesc
            n_blocks += 1
            for f2, _, v2 in _fields(v):
                if f2 == 3:
                    var = _parse_var(v2)
                    vars_[var.name] = var
                elif f2 == 4:
                    ops.append(_parse_op(v2))
        elif fno == 4:
            for f2, _, v2 in _fields(v):
                if f2 == 1:
                    version = _s64(v2)
    if n_blocks != 1:
        raise ValueError(f"expected a single-block program, found {n_blocks}")
    return Program(vars=vars_, ops=ops, version=version)


# --------------------------------------------------------------------------- #
# parameters
# --------------------------------------------------------------------------- #


def read_params_bytes(model_dir: str) -> bytes:
    """Return the raw ``.pdiparams`` byte stream, joining split files in manifest order."""
    single = os.path.join(model_dir, "inference.pdiparams")
    if os.path.exists(single):
        with open(single, "rb") as f:
            return f.read()
    manifest = os.path.join(model_dir, "fs_manifest.csv")
    if not os.path.exists(manifest):
        raise FileNotFoundError(f"no inference.pdiparams or fs_manifest.csv in {model_dir}")
    parts: List[str] = []
    with open(manifest, newline="") as f:
        for row in csv.reader(f):
            if not row:
                continue
            name = row[0].strip()
            if name.lower() in ("filename", "name") or not name:
                continue
            parts.append(name)
    chunks = []
    for name in parts:
        with open(os.path.join(model_dir, os.path.basename(name)), "rb") as f:
            chunks.append(f.read())
    return b"".join(chunks)


def parse_params(buf: bytes, names: List[str]) -> Dict[str, np.ndarray]:
    """Decode the record stream; ``names`` = persistable var names (any order)."""
    out: Dict[str, np.ndarray] = {}
    pos = 0
    for name in sorted(names):
        lod_version, = struct.unpack_from("<I", buf, pos)
        pos += 4
        lod_levels, = struct.unpack_from("<Q", buf, pos)
        pos += 8
        for _ in range(lod_levels):
            sz, = struct.unpack_from("<Q", buf, pos)
            pos += 8 + sz
        tensor_version, = struct.unpack_from("<I", buf, pos)
        pos += 4
        desc_len, = struct.unpack_from("<i", buf, pos)
        pos += 4
        dtype, dims = _parse_tensor_desc(buf[pos:pos + desc_len])
        pos += desc_len
        np_dtype = np.dtype(_DTYPES[dtype])
        count = int(np.prod(dims)) if dims else 1
        nbytes = count * np_dtype.itemsize
        arr = np.frombuffer(buf, dtype=np_dtype, count=count, offset=pos).reshape(dims)
        pos += nbytes
        out[name] = arr
        if lod_version != 0 or tensor_version != 0:
            raise ValueError(f"unexpected record version for {name}")
    if pos != len(buf):
        raise ValueError(f"parameter stream has {len(buf) - pos} trailing bytes")
    return out


@dataclass
class Model:
    """One shipped inference model: graph + fp32 parameters."""
    program: Program
    params: Dict[str, np.ndarray]
    model_dir: str = ""


def load_model(model_dir: str) -> Model:
    with open(os.path.join(model_dir, "inference.pdmodel"), "rb") as f:
        program = parse_program(f.read())
    names = [v.name for v in program.vars.values() if v.persistable and v.name not in ("feed", "fetch")]
    params = parse_params(read_params_bytes(model_dir), names)
    for n in names:
        want = program.vars[n].dims
        if list(params[n].shape) != list(want):
            raise ValueError(f"{n}: param shape {params[n].shape} != declared {want}")
    return Model(program=program, params=params, model_dir=model_dir)
