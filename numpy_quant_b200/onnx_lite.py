"""Dependency-free ONNX container: wire-format reader/writer + plain-Python protos.

`Model.from_onnx` (reference `numpy_quant/model.py:249-292`) only touches a
handful of attributes of an `onnx.ModelProto`:

    graph.initializer[*] -> tensor (name, dims, payload)
    graph.input[*].name, graph.output[*].name
    graph.node[*].{name, op_type, attribute, input, output}

This module provides light objects with exactly those attributes, a protobuf
wire-format decoder so committed ``.onnx`` files load without the `onnx`
package, an encoder so fixtures can be written, and the two helper functions the
reference calls (`to_array`, `get_attribute_value`).  Field numbers follow the
public ONNX schema (onnx.proto3).
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field
from typing import Any, Iterable

import numpy as np

# TensorProto.DataType (subset)
FLOAT, UINT8, INT8, INT32, INT64, BOOL, DOUBLE = 1, 2, 3, 6, 7, 9, 11
_NP_OF = {FLOAT: np.float32, UINT8: np.uint8, INT8: np.int8, INT32: np.int32,
          INT64: np.int64, BOOL: np.bool_, DOUBLE: np.float64}
_DT_OF = {np.dtype(v): k for k, v in _NP_OF.items()}

# AttributeProto.AttributeType (subset)
AT_FLOAT, AT_INT, AT_STRING, AT_TENSOR, AT_FLOATS, AT_INTS, AT_STRINGS = 1, 2, 3, 4, 6, 7, 8


@dataclass
class TensorProto:
    name: str = ""
    dims: list[int] = field(default_factory=list)
    data_type: int = FLOAT
    raw_data: bytes = b""
    float_data: list[float] = field(default_factory=list)
    int64_data: list[int] = field(default_factory=list)
    int32_data: list[int] = field(default_factory=list)
    data_location: int = 0          # 1 == EXTERNAL (payload not in the file)
    external_data: dict[str, str] = field(default_factory=dict)

    FLOAT = FLOAT
    INT64 = INT64


@dataclass
class AttributeProto:
    name: str = ""
    type: int = 0
    f: float = 0.0
    i: int = 0
    s: bytes = b""
    t: TensorProto | None = None
    floats: list[float] = field(default_factory=list)
    ints: list[int] = field(default_factory=list)
    strings: list[bytes] = field(default_factory=list)


@dataclass
class NodeProto:
    name: str = ""
    op_type: str = ""
    input: list[str] = field(default_factory=list)
    output: list[str] = field(default_factory=list)
    attribute: list[AttributeProto] = field(default_factory=list)


@dataclass
class ValueInfoProto:
    name: str = ""
    elem_type: int = FLOAT
    shape: list[Any] = field(default_factory=list)   # ints or symbolic names


@dataclass
class GraphProto:
    name: str = ""
    node: list[NodeProto] = field(default_factory=list)
    initializer: list[TensorProto] = field(default_factory=list)
    input: list[ValueInfoProto] = field(default_factory=list)
    output: list[ValueInfoProto] = field(default_factory=list)


@dataclass
class ModelProto:
    ir_version: int = 8
    producer_name: str = "numpy_quant_b200"
    opset: int = 17
    graph: GraphProto = field(default_factory=GraphProto)


# --------------------------------------------------------------------------
# helpers mirroring onnx.numpy_helper.to_array / onnx.helper.get_attribute_value
# --------------------------------------------------------------------------
def to_array(t: TensorProto) -> np.ndarray:
    if t.data_location == 1 and not (t.raw_data or t.float_data or t.int64_data):
        raise ValueError(f"tensor {t.name!r} stores its payload externally "
                         f"({t.external_data.get('location', '?')}); not available")
    dt = _NP_OF[t.data_type]
    shape = tuple(int(d) for d in t.dims)
    if t.raw_data:
        arr = np.frombuffer(t.raw_data, dtype=np.dtype(dt).newbyteorder("<")).astype(dt)
    elif t.data_type == FLOAT:
        arr = np.asarray(t.float_data, dtype=np.float32)
    elif t.data_type == INT64:
        arr = np.asarray(t.int64_data, dtype=np.int64)
    elif t.data_type in (INT32, INT8, UINT8, BOOL):
        arr = np.asarray(t.int32_data, dtype=dt)
    else:
        arr = np.zeros(0, dtype=dt)
    return arr.reshape(shape).copy()


def from_array(arr: np.ndarray, name: str = "") -> TensorProto:
    arr = np.asarray(arr)
    return TensorProto(name=name, dims=list(arr.shape), data_type=_DT_OF[arr.dtype],
                       raw_data=np.ascontiguousarray(arr).astype(arr.dtype.newbyteorder("<")).tobytes())


def get_attribute_value(a: AttributeProto):
    if a.type == AT_FLOAT:
        return a.f
    if a.type == AT_INT:
        return a.i
    if a.type == AT_STRING:
        return a.s
    if a.type == AT_TENSOR:
        return a.t
    if a.type == AT_FLOATS:
        return list(a.floats)
    if a.type == AT_INTS:
        return list(a.ints)
    if a.type == AT_STRINGS:
        return list(a.strings)
    raise ValueError(f"attribute {a.name!r}: unsupported type {a.type}")


def make_attribute(name: str, value) -> AttributeProto:
    if isinstance(value, TensorProto):
        return AttributeProto(name=name, type=AT_TENSOR, t=value)
    if isinstance(value, np.ndarray):
        return AttributeProto(name=name, type=AT_TENSOR, t=from_array(value))
    if isinstance(value, (bool, int, np.integer)):
        return AttributeProto(name=name, type=AT_INT, i=int(value))
    if isinstance(value, (float, np.floating)):
        return AttributeProto(name=name, type=AT_FLOAT, f=float(value))
    if isinstance(value, (bytes, str)):
        return AttributeProto(name=name, type=AT_STRING, s=value.encode() if isinstance(value, str) else value)
    if isinstance(value, (list, tuple)):
        if all(isinstance(v, (int, np.integer)) for v in value):
            return AttributeProto(name=name, type=AT_INTS, ints=[int(v) for v in value])
        return AttributeProto(name=name, type=AT_FLOATS, floats=[float(v) for v in value])
    raise TypeError(f"attribute {name!r}: cannot encode {type(value)}")


def make_node(op_type: str, inputs: Iterable[str], outputs: Iterable[str], name: str = "", **attrs) -> NodeProto:
    return NodeProto(name=name, op_type=op_type, input=list(inputs), output=list(outputs),
                     attribute=[make_attribute(k, v) for k, v in attrs.items()])


# --------------------------------------------------------------------------
# protobuf wire format
# --------------------------------------------------------------------------
def _rd_varint(buf: memoryview, pos: int) -> tuple[int, int]:
    out = shift = 0
    while True:
        b = buf[pos]
        pos += 1
        out |= (b & 0x7F) << shift
        if b < 0x80:
            return out, pos
        shift += 7


def _signed64(u: int) -> int:
    return u - (1 << 64) if u >= (1 << 63) else u


def _fields(buf: memoryview):
    """Yield (field_number, wire_type, value) triples of one message."""
    pos, n = 0, len(buf)
    while pos < n:
        key, pos = _rd_varint(buf, pos)
        fno, wt = key >> 3, key & 7
        if wt == 0:
            val, pos = _rd_varint(buf, pos)
        elif wt == 1:
            val, pos = bytes(buf[pos:pos + 8]), pos + 8
        elif wt == 2:
            ln, pos = _rd_varint(buf, pos)
            val, pos = buf[pos:pos + ln], pos + ln
        elif wt == 5:
            val, pos = bytes(buf[pos:pos + 4]), pos + 4
        else:
            raise ValueError(f"unsupported protobuf wire type {wt}")
        yield fno, wt, val


def _packed_varints(wt, val, signed=True) -> list[int]:
    if wt == 0:
        return [_signed64(val) if signed else val]
    out, pos = [], 0
    while pos < len(val):
        v, pos = _rd_varint(val, pos)
        out.append(_signed64(v) if signed else v)
    return out


def _packed_f32(wt, val) -> list[float]:
    if wt == 5:
        return [struct.unpack("<f", val)[0]]
    return list(np.frombuffer(bytes(val), dtype="<f4"))


def _parse_tensor(buf) -> TensorProto:
    t = TensorProto()
    for fno, wt, v in _fields(buf):
        if fno == 1:
            t.dims += _packed_varints(wt, v)
        elif fno == 2:
            t.data_type = v
        elif fno == 4:
            t.float_data += _packed_f32(wt, v)
        elif fno == 5:
            t.int32_data += _packed_varints(wt, v)
        elif fno == 7:
            t.int64_data += _packed_varints(wt, v)
        elif fno == 8:
            t.name = bytes(v).decode()
        elif fno == 9:
            t.raw_data = bytes(v)
        elif fno == 13:
            kv = {f: bytes(x).decode() for f, _, x in _fields(v)}
            t.external_data[kv.get(1, "")] = kv.get(2, "")
        elif fno == 14:
            t.data_location = v
    return t


def _parse_attribute(buf) -> AttributeProto:
    a = AttributeProto()
    for fno, wt, v in _fields(buf):
        if fno == 1:
            a.name = bytes(v).decode()
        elif fno == 2:
            a.f = struct.unpack("<f", v)[0]
        elif fno == 3:
            a.i = _signed64(v)
        elif fno == 4:
            a.s = bytes(v)
        elif fno == 5:
            a.t = _parse_tensor(v)
        elif fno == 7:
            a.floats += _packed_f32(wt, v)
        elif fno == 8:
            a.ints += _packed_varints(wt, v)
        elif fno == 9:
            a.strings.append(bytes(v))
        elif fno == 20:
            a.type = v
    if a.type == 0:  # old exporters omit `type`; infer it
        a.type = (AT_TENSOR if a.t is not None else AT_INTS if a.ints else AT_FLOATS if a.floats
                  else AT_STRING if a.s else AT_INT)
    return a


def _parse_node(buf) -> NodeProto:
    n = NodeProto()
    for fno, _, v in _fields(buf):
        if fno == 1:
            n.input.append(bytes(v).decode())
        elif fno == 2:
            n.output.append(bytes(v).decode())
        elif fno == 3:
            n.name = bytes(v).decode()
        elif fno == 4:
            n.op_type = bytes(v).decode()
        elif fno == 5:
            n.attribute.append(_parse_attribute(v))
    return n


def _parse_value_info(buf) -> ValueInfoProto:
    vi = ValueInfoProto()
    for fno, _, v in _fields(buf):
        if fno == 1:
            vi.name = bytes(v).decode()
        elif fno == 2:                                   # TypeProto
            for f2, _, v2 in _fields(v):
                if f2 != 1:                              # tensor_type
                    continue
                for f3, _, v3 in _fields(v2):
                    if f3 == 1:
                        vi.elem_type = v3
                    elif f3 == 2:                        # TensorShapeProto
                        for f4, _, v4 in _fields(v3):
                            if f4 != 1:
                                continue
                            dim: Any = None
                            for f5, _, v5 in _fields(v4):
                                if f5 == 1:
                                    dim = _signed64(v5)
                                elif f5 == 2:
                                    dim = bytes(v5).decode()
                            vi.shape.append(dim)
    return vi


def _parse_graph(buf) -> GraphProto:
    g = GraphProto()
    for fno, _, v in _fields(buf):
        if fno == 1:
            g.node.append(_parse_node(v))
        elif fno == 2:
            g.name = bytes(v).decode()
        elif fno == 5:
            g.initializer.append(_parse_tensor(v))
        elif fno == 11:
            g.input.append(_parse_value_info(v))
        elif fno == 12:
            g.output.append(_parse_value_info(v))
    return g


def load_model_from_string(data: bytes) -> ModelProto:
    m = ModelProto()
    for fno, _, v in _fields(memoryview(data)):
        if fno == 1:
            m.ir_version = v
        elif fno == 2:
            m.producer_name = bytes(v).decode()
        elif fno == 7:
            m.graph = _parse_graph(v)
        elif fno == 8:
            for f2, _, v2 in _fields(v):
                if f2 == 2:
                    m.opset = v2
    return m


def load_external_data(m: ModelProto, base_dir) -> ModelProto:
    """Pull the payload of every EXTERNAL initializer (onnx external-data convention, as written by the reference's
    exporter `models/vit.py:71-86` with save_as_external_data=True) into `raw_data`: `location` is a path relative
    to the model file, `offset` / `length` select the byte range (whole file when absent)."""
    import os
    cache: dict[str, memoryview] = {}
    for t in m.graph.initializer:
        if t.data_location != 1 or t.raw_data:
            continue
        loc = t.external_data.get("location")
        if not loc:
            raise ValueError(f"tensor {t.name!r} is EXTERNAL but names no location")
        # containment: `location` comes from the (untrusted) model file -- it must name a regular file inside the
        # model's own directory (no absolute paths, no `..`, no symlinks leading elsewhere)
        base = os.path.realpath(str(base_dir))
        if os.path.isabs(loc) or loc.startswith(("/", "\\")):
            raise ValueError(f"tensor {t.name!r}: absolute external-data location {loc!r} refused")
        joined = os.path.join(base, loc)
        full = os.path.realpath(joined)
        if os.path.commonpath([base, full]) != base or os.path.islink(joined):
            raise ValueError(f"tensor {t.name!r}: external-data location {loc!r} escapes the model directory")
        if full not in cache:
            with open(full, "rb") as fh:
                cache[full] = memoryview(fh.read())
        blob = cache[full]
        off = int(t.external_data.get("offset", "0") or 0)
        n = t.external_data.get("length")
        n = int(n) if n not in (None, "") else len(blob) - off
        if off < 0 or n < 0 or off + n > len(blob):
            raise ValueError(f"tensor {t.name!r}: external range [{off}, {off + n}) outside {loc} ({len(blob)} bytes)")
        t.raw_data = bytes(blob[off:off + n])
        t.data_location = 0
        t.external_data = {}
    return m


def load(path, load_external: bool = True) -> ModelProto:
    import os
    with open(path, "rb") as fh:
        m = load_model_from_string(fh.read())
    if load_external:
        load_external_data(m, os.path.dirname(os.path.abspath(str(path))))
    return m


# ---- encoder ---------------------------------------------------------------
def _varint(v: int) -> bytes:
    v &= (1 << 64) - 1
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        out.append(b | (0x80 if v else 0))
        if not v:
            return bytes(out)


def _key(fno: int, wt: int) -> bytes:
    return _varint((fno << 3) | wt)


def _ld(fno: int, payload: bytes) -> bytes:
    return _key(fno, 2) + _varint(len(payload)) + payload


def _enc_tensor(t: TensorProto) -> bytes:
    out = b"".join(_key(1, 0) + _varint(d) for d in t.dims)
    out += _key(2, 0) + _varint(t.data_type)
    if t.float_data:
        out += _ld(4, np.asarray(t.float_data, "<f4").tobytes())
    if t.int64_data:
        out += _ld(7, b"".join(_varint(v) for v in t.int64_data))
    if t.name:
        out += _ld(8, t.name.encode())
    if t.raw_data and t.data_location != 1:
        out += _ld(9, t.raw_data)
    if t.data_location == 1:
        for k, v in t.external_data.items():
            out += _ld(13, _ld(1, k.encode()) + _ld(2, str(v).encode()))
        out += _key(14, 0) + _varint(1)
    return out


def _enc_attribute(a: AttributeProto) -> bytes:
    out = _ld(1, a.name.encode())
    if a.type == AT_FLOAT:
        out += _key(2, 5) + struct.pack("<f", a.f)
    elif a.type == AT_INT:
        out += _key(3, 0) + _varint(a.i)
    elif a.type == AT_STRING:
        out += _ld(4, a.s)
    elif a.type == AT_TENSOR:
        out += _ld(5, _enc_tensor(a.t))
    elif a.type == AT_FLOATS:
        out += _ld(7, np.asarray(a.floats, "<f4").tobytes())
    elif a.type == AT_INTS:
        out += _ld(8, b"".join(_varint(v) for v in a.ints))
    out += _key(20, 0) + _varint(a.type)
    return out


def _enc_node(n: NodeProto) -> bytes:
    out = b"".join(_ld(1, s.encode()) for s in n.input)
    out += b"".join(_ld(2, s.encode()) for s in n.output)
    out += _ld(3, n.name.encode()) + _ld(4, n.op_type.encode())
    out += b"".join(_ld(5, _enc_attribute(a)) for a in n.attribute)
    return out


def _enc_value_info(vi: ValueInfoProto) -> bytes:
    dims = b""
    for d in vi.shape:
        dim = _ld(2, d.encode()) if isinstance(d, str) else _key(1, 0) + _varint(int(d))
        dims += _ld(1, dim)
    tensor_type = _key(1, 0) + _varint(vi.elem_type) + _ld(2, dims)
    return _ld(1, vi.name.encode()) + _ld(2, _ld(1, tensor_type))


def serialize(m: ModelProto) -> bytes:
    g = m.graph
    gb = b"".join(_ld(1, _enc_node(n)) for n in g.node)
    gb += _ld(2, g.name.encode())
    gb += b"".join(_ld(5, _enc_tensor(t)) for t in g.initializer)
    gb += b"".join(_ld(11, _enc_value_info(v)) for v in g.input)
    gb += b"".join(_ld(12, _enc_value_info(v)) for v in g.output)
    out = _key(1, 0) + _varint(m.ir_version) + _ld(2, m.producer_name.encode()) + _ld(7, gb)
    out += _ld(8, _ld(1, b"") + _key(2, 0) + _varint(m.opset))
    return out


def save(m: ModelProto, path, save_as_external_data: bool = False, location: str | None = None,
         size_threshold: int = 1024) -> None:
    """Write the model; with save_as_external_data the raw payload of every initializer of at least `size_threshold`
    bytes goes to one side file `location` (default: <model file name>.data) next to the model, 64-byte aligned,
    referenced by location / offset / length entries -- the layout onnx.save(..., save_as_external_data=True) produces."""
    import copy
    import os
    path = str(path)
    if not save_as_external_data:
        with open(path, "wb") as fh:
            fh.write(serialize(m))
        return
    location = location or os.path.basename(path) + ".data"
    out = copy.copy(m)
    out.graph = copy.copy(m.graph)
    out.graph.initializer = []
    blob = bytearray()
    for t in m.graph.initializer:
        payload = t.raw_data or to_array(t).astype(np.dtype(_NP_OF[t.data_type]).newbyteorder("<")).tobytes()
        if len(payload) < size_threshold:
            out.graph.initializer.append(t)
            continue
        blob.extend(b"\0" * (-len(blob) % 64))
        ext = TensorProto(name=t.name, dims=list(t.dims), data_type=t.data_type, data_location=1,
                          external_data={"location": location, "offset": str(len(blob)), "length": str(len(payload))})
        blob.extend(payload)
        out.graph.initializer.append(ext)
    with open(os.path.join(os.path.dirname(os.path.abspath(path)), location), "wb") as fh:
        fh.write(bytes(blob))
    with open(path, "wb") as fh:
        fh.write(serialize(out))
