"""Dependency-free reader/writer for the subset of the ONNX protobuf format floodsr needs.

The reference hands `model_infer.onnx` to ONNX Runtime (`floodsr/engine/ort.py:54`); this
backend instead parses the file itself to get at the graph nodes and the initializers (the network
weights), because neither `onnx` nor `onnxruntime` is a dependency of the B200 engine.

Only protobuf wire-format primitives are implemented (varint, 64-bit, length-delimited, 32-bit) and
only the ONNX messages/fields the hot path uses: ModelProto, GraphProto, NodeProto, AttributeProto,
TensorProto, ValueInfoProto.  Field numbers follow onnx.proto (IR version 7 era).

The writer exists so that a random-init network of the same architecture can be produced offline as
a real `.onnx` file (see `floodsr_b200/h1.py`) and pushed through exactly the same load path.
"""

from __future__ import annotations

import struct
from dataclasses import dataclass, field
from pathlib import Path
from typing import Any, Iterator

import numpy as np

# --------------------------------------------------------------------------------------
# wire primitives
# --------------------------------------------------------------------------------------

_WT_VARINT, _WT_I64, _WT_LEN, _WT_I32 = 0, 1, 2, 5


def _read_varint(buf: memoryview, pos: int) -> tuple[int, int]:
    result = 0
    shift = 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not (b & 0x80):
            return result, pos
        shift += 7
        if shift > 70:
            raise ValueError("malformed varint in ONNX file")


def _iter_fields(buf: memoryview) -> Iterator[tuple[int, int, Any]]:
    """Yield (field_number, wire_type, value) for one message body."""
    pos = 0
    end = len(buf)
    while pos < end:
        key, pos = _read_varint(buf, pos)
        fnum, wt = key >> 3, key & 7
        if wt == _WT_VARINT:
            val, pos = _read_varint(buf, pos)
        elif wt == _WT_I64:
            val = bytes(buf[pos : pos + 8])
            pos += 8
        elif wt == _WT_LEN:
            ln, pos = _read_varint(buf, pos)
            val = buf[pos : pos + ln]
            pos += ln
        elif wt == _WT_I32:
            val = bytes(buf[pos : pos + 4])
            pos += 4
        else:
            raise ValueError(f"unsupported protobuf wire type {wt} (field {fnum})")
        if pos > end:
            raise ValueError("truncated ONNX file")
        yield fnum, wt, val


def _signed64(v: int) -> int:
    return v - (1 << 64) if v >= (1 << 63) else v


def _packed_varints(val: Any, wt: int) -> list[int]:
    if wt == _WT_VARINT:
        return [_signed64(val)]
    out = []
    pos = 0
    mv = val
    while pos < len(mv):
        v, pos = _read_varint(mv, pos)
        out.append(_signed64(v))
    return out


def _enc_varint(v: int) -> bytes:
    if v < 0:
        v += 1 << 64
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _enc_key(fnum: int, wt: int) -> bytes:
    return _enc_varint((fnum << 3) | wt)


def _enc_len(fnum: int, payload: bytes) -> bytes:
    return _enc_key(fnum, _WT_LEN) + _enc_varint(len(payload)) + payload


def _enc_int(fnum: int, v: int) -> bytes:
    return _enc_key(fnum, _WT_VARINT) + _enc_varint(int(v))


def _enc_str(fnum: int, s: str) -> bytes:
    return _enc_len(fnum, s.encode("utf-8"))


# --------------------------------------------------------------------------------------
# data model
# --------------------------------------------------------------------------------------

# TensorProto.DataType
DT_FLOAT, DT_UINT8, DT_INT8, DT_INT32, DT_INT64, DT_BOOL, DT_FLOAT16, DT_DOUBLE = 1, 2, 3, 6, 7, 9, 10, 11
_DT_NP = {
    DT_FLOAT: np.float32,
    DT_UINT8: np.uint8,
    DT_INT8: np.int8,
    DT_INT32: np.int32,
    DT_INT64: np.int64,
    DT_BOOL: np.bool_,
    DT_FLOAT16: np.float16,
    DT_DOUBLE: np.float64,
}
_NP_DT = {np.dtype(v): k for k, v in _DT_NP.items()}

# AttributeProto.AttributeType
_AT_FLOAT, _AT_INT, _AT_STRING, _AT_TENSOR, _AT_FLOATS, _AT_INTS, _AT_STRINGS = 1, 2, 3, 4, 6, 7, 8


@dataclass
class OnnxNode:
    op_type: str
    inputs: list[str]
    outputs: list[str]
    name: str = ""
    attrs: dict[str, Any] = field(default_factory=dict)


@dataclass
class OnnxValueInfo:
    name: str
    elem_type: int
    shape: list[int | str | None]


@dataclass
class OnnxModel:
    nodes: list[OnnxNode]
    initializers: dict[str, np.ndarray]
    inputs: list[OnnxValueInfo]
    outputs: list[OnnxValueInfo]
    ir_version: int = 7
    opset: int = 13
    producer_name: str = ""
    producer_version: str = ""
    graph_name: str = "graph"

    def initializer_param_count(self) -> int:
        """Same count the reference notebook prints (`infer_test_tiles.ipynb` cell 9)."""
        return int(sum(int(np.prod(a.shape)) for a in self.initializers.values()))


# --------------------------------------------------------------------------------------
# reader
# --------------------------------------------------------------------------------------


def _parse_tensor(buf: memoryview) -> tuple[str, np.ndarray]:
    dims: list[int] = []
    dtype = DT_FLOAT
    name = ""
    raw = None
    float_data: list[float] = []
    int32_data: list[int] = []
    int64_data: list[int] = []
    double_data: list[float] = []
    for fnum, wt, val in _iter_fields(buf):
        if fnum == 1:
            dims.extend(_packed_varints(val, wt))
        elif fnum == 2:
            dtype = int(val)
        elif fnum == 4:
            if wt == _WT_LEN:
                float_data.extend(np.frombuffer(bytes(val), dtype="<f4").tolist())
            else:
                float_data.append(struct.unpack("<f", val)[0])
        elif fnum == 5:
            int32_data.extend(_packed_varints(val, wt))
        elif fnum == 7:
            int64_data.extend(_packed_varints(val, wt))
        elif fnum == 8:
            name = bytes(val).decode("utf-8")
        elif fnum == 9:
            raw = bytes(val)
        elif fnum == 10:
            if wt == _WT_LEN:
                double_data.extend(np.frombuffer(bytes(val), dtype="<f8").tolist())
            else:
                double_data.append(struct.unpack("<d", val)[0])
        elif fnum in (13, 14):
            if fnum == 14 and int(val) != 0:
                raise ValueError(f"tensor '{name}': external data is not supported")
    if dtype not in _DT_NP:
        raise ValueError(f"tensor '{name}': unsupported ONNX data type {dtype}")
    np_dt = np.dtype(_DT_NP[dtype])
    if raw is not None:
        arr = np.frombuffer(raw, dtype=np_dt.newbyteorder("<")).astype(np_dt, copy=True)
    elif dtype == DT_FLOAT:
        arr = np.asarray(float_data, dtype=np.float32)
    elif dtype == DT_DOUBLE:
        arr = np.asarray(double_data, dtype=np.float64)
    elif dtype == DT_INT64:
        arr = np.asarray(int64_data, dtype=np.int64)
    elif dtype == DT_FLOAT16:
        arr = np.asarray(int32_data, dtype=np.uint16).view(np.float16)
    else:
        arr = np.asarray(int32_data).astype(np_dt)
    n = int(np.prod(dims)) if dims else arr.size
    if arr.size != n:
        raise ValueError(f"tensor '{name}': {arr.size} elements for dims {dims}")
    return name, arr.reshape(dims)


def _parse_attribute(buf: memoryview) -> tuple[str, Any]:
    name = ""
    f = i = s = t = None
    floats: list[float] = []
    ints: list[int] = []
    strings: list[bytes] = []
    atype = 0
    for fnum, wt, val in _iter_fields(buf):
        if fnum == 1:
            name = bytes(val).decode("utf-8")
        elif fnum == 2:
            f = struct.unpack("<f", val)[0]
        elif fnum == 3:
            i = _signed64(val)
        elif fnum == 4:
            s = bytes(val)
        elif fnum == 5:
            t = _parse_tensor(val)[1]
        elif fnum == 7:
            if wt == _WT_LEN:
                floats.extend(np.frombuffer(bytes(val), dtype="<f4").tolist())
            else:
                floats.append(struct.unpack("<f", val)[0])
        elif fnum == 8:
            ints.extend(_packed_varints(val, wt))
        elif fnum == 9:
            strings.append(bytes(val))
        elif fnum == 20:
            atype = int(val)
    if atype == _AT_FLOAT or (atype == 0 and f is not None):
        return name, float(f if f is not None else 0.0)
    if atype == _AT_INT or (atype == 0 and i is not None):
        return name, int(i if i is not None else 0)
    if atype == _AT_STRING or (atype == 0 and s is not None):
        return name, (s or b"").decode("utf-8", errors="replace")
    if atype == _AT_TENSOR or (atype == 0 and t is not None):
        return name, t
    if atype == _AT_FLOATS or (atype == 0 and floats):
        return name, [float(x) for x in floats]
    if atype == _AT_INTS or (atype == 0 and ints):
        return name, [int(x) for x in ints]
    if atype == _AT_STRINGS or (atype == 0 and strings):
        return name, [x.decode("utf-8", errors="replace") for x in strings]
    return name, None


def _parse_node(buf: memoryview) -> OnnxNode:
    node = OnnxNode(op_type="", inputs=[], outputs=[])
    for fnum, _wt, val in _iter_fields(buf):
        if fnum == 1:
            node.inputs.append(bytes(val).decode("utf-8"))
        elif fnum == 2:
            node.outputs.append(bytes(val).decode("utf-8"))
        elif fnum == 3:
            node.name = bytes(val).decode("utf-8")
        elif fnum == 4:
            node.op_type = bytes(val).decode("utf-8")
        elif fnum == 5:
            k, v = _parse_attribute(val)
            node.attrs[k] = v
    return node


def _parse_value_info(buf: memoryview) -> OnnxValueInfo:
    name = ""
    elem_type = 0
    shape: list[int | str | None] = []
    for fnum, _wt, val in _iter_fields(buf):
        if fnum == 1:
            name = bytes(val).decode("utf-8")
        elif fnum == 2:  # TypeProto
            for f2, _w2, v2 in _iter_fields(val):
                if f2 != 1:  # tensor_type
                    continue
                for f3, _w3, v3 in _iter_fields(v2):
                    if f3 == 1:
                        elem_type = int(v3)
                    elif f3 == 2:  # TensorShapeProto
                        for f4, _w4, v4 in _iter_fields(v3):
                            if f4 != 1:
                                continue
                            dim: int | str | None = None
                            for f5, _w5, v5 in _iter_fields(v4):
                                if f5 == 1:
                                    dim = _signed64(v5)
                                elif f5 == 2:
                                    dim = bytes(v5).decode("utf-8")
                            shape.append(dim)
    return OnnxValueInfo(name=name, elem_type=elem_type, shape=shape)


def load_onnx(path: str | Path) -> OnnxModel:
    """Parse an ONNX file into nodes, initializers and graph inputs/outputs."""
    data = Path(path).read_bytes()
    if len(data) < 16:
        raise ValueError(f"{path}: too small to be an ONNX model ({len(data)} bytes)")
    buf = memoryview(data)
    model = OnnxModel(nodes=[], initializers={}, inputs=[], outputs=[])
    graph_buf = None
    try:
        for fnum, _wt, val in _iter_fields(buf):
            if fnum == 1:
                model.ir_version = int(val)
            elif fnum == 2:
                model.producer_name = bytes(val).decode("utf-8")
            elif fnum == 3:
                model.producer_version = bytes(val).decode("utf-8")
            elif fnum == 7:
                graph_buf = val
            elif fnum == 8:
                domain, version = "", 0
                for f2, _w2, v2 in _iter_fields(val):
                    if f2 == 1:
                        domain = bytes(v2).decode("utf-8")
                    elif f2 == 2:
                        version = int(v2)
                if domain in ("", "ai.onnx"):
                    model.opset = version
    except (IndexError, struct.error) as exc:
        raise ValueError(f"{path}: not a valid ONNX protobuf ({exc})") from exc
    if graph_buf is None:
        raise ValueError(f"{path}: ONNX model has no graph")
    for fnum, _wt, val in _iter_fields(graph_buf):
        if fnum == 1:
            model.nodes.append(_parse_node(val))
        elif fnum == 2:
            model.graph_name = bytes(val).decode("utf-8")
        elif fnum == 5:
            name, arr = _parse_tensor(val)
            model.initializers[name] = arr
        elif fnum == 11:
            model.inputs.append(_parse_value_info(val))
        elif fnum == 12:
            model.outputs.append(_parse_value_info(val))
    # graph inputs that are initializers are weights, not feeds (IR < 4 style files list both)
    model.inputs = [vi for vi in model.inputs if vi.name not in model.initializers]
    return model


# --------------------------------------------------------------------------------------
# writer
# --------------------------------------------------------------------------------------


def _enc_tensor(name: str, arr: np.ndarray) -> bytes:
    arr = np.ascontiguousarray(arr)
    if arr.dtype not in _NP_DT:
        raise ValueError(f"cannot serialise dtype {arr.dtype}")
    out = bytearray()
    for d in arr.shape:
        out += _enc_int(1, d)
    out += _enc_int(2, _NP_DT[arr.dtype])
    out += _enc_str(8, name)
    out += _enc_len(9, arr.astype(arr.dtype.newbyteorder("<"), copy=False).tobytes())
    return bytes(out)


def _enc_attribute(name: str, value: Any) -> bytes:
    out = bytearray(_enc_str(1, name))
    if isinstance(value, bool):
        value = int(value)
    if isinstance(value, float):
        out += _enc_key(2, _WT_I32) + struct.pack("<f", value)
        out += _enc_int(20, _AT_FLOAT)
    elif isinstance(value, (int, np.integer)):
        out += _enc_int(3, int(value))
        out += _enc_int(20, _AT_INT)
    elif isinstance(value, str):
        out += _enc_len(4, value.encode("utf-8"))
        out += _enc_int(20, _AT_STRING)
    elif isinstance(value, np.ndarray):
        out += _enc_len(5, _enc_tensor("", value))
        out += _enc_int(20, _AT_TENSOR)
    elif isinstance(value, (list, tuple)) and value and all(isinstance(v, float) for v in value):
        for v in value:
            out += _enc_key(7, _WT_I32) + struct.pack("<f", v)
        out += _enc_int(20, _AT_FLOATS)
    elif isinstance(value, (list, tuple)):
        for v in value:
            out += _enc_int(8, int(v))
        out += _enc_int(20, _AT_INTS)
    else:
        raise ValueError(f"cannot serialise attribute {name}={value!r}")
    return bytes(out)


def _enc_node(node: OnnxNode) -> bytes:
    out = bytearray()
    for s in node.inputs:
        out += _enc_str(1, s)
    for s in node.outputs:
        out += _enc_str(2, s)
    if node.name:
        out += _enc_str(3, node.name)
    out += _enc_str(4, node.op_type)
    for k, v in node.attrs.items():
        out += _enc_len(5, _enc_attribute(k, v))
    return bytes(out)


def _enc_value_info(vi: OnnxValueInfo) -> bytes:
    shape = bytearray()
    for d in vi.shape:
        if isinstance(d, str):
            dim = _enc_str(2, d)
        elif d is None:
            dim = b""
        else:
            dim = _enc_int(1, d)
        shape += _enc_len(1, dim)
    tensor_type = _enc_int(1, vi.elem_type) + _enc_len(2, bytes(shape))
    type_proto = _enc_len(1, tensor_type)
    return _enc_str(1, vi.name) + _enc_len(2, type_proto)


def save_onnx(model: OnnxModel, path: str | Path) -> None:
    """Serialise an OnnxModel (raw_data initializers) to `path`."""
    graph = bytearray()
    for node in model.nodes:
        graph += _enc_len(1, _enc_node(node))
    graph += _enc_str(2, model.graph_name)
    for name, arr in model.initializers.items():
        graph += _enc_len(5, _enc_tensor(name, arr))
    for vi in model.inputs:
        graph += _enc_len(11, _enc_value_info(vi))
    for vi in model.outputs:
        graph += _enc_len(12, _enc_value_info(vi))
    out = bytearray()
    out += _enc_int(1, model.ir_version)
    out += _enc_str(2, model.producer_name)
    out += _enc_str(3, model.producer_version)
    out += _enc_len(7, bytes(graph))
    out += _enc_len(8, _enc_str(1, "") + _enc_int(2, model.opset))
    Path(path).write_bytes(bytes(out))
