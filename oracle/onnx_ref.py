"""Oracle for the network forward pass (TEST INFRASTRUCTURE ONLY): ONNX graph executed with torch CPU ops.

Stands in for `onnxruntime.InferenceSession.run` at `/root/reference/floodsr/engine/ort.py:193`
(third-party dependency `onnxruntime>=1.18`, `pyproject.toml:15`; pinned 1.24.2 in
`container/bookworm/pip-freeze.base.txt:10`), which is absent offline.  Each ONNX operator is evaluated
by the torch CPU kernel with the same published semantics (ONNX opset 13 operator spec).

The file decoder here is deliberately independent from the product's `floodsr_b200/onnx_io.py`
(a schema-free protobuf walk + a table of field numbers), so a bug in one reader shows up as a
disagreement between the two.

PARITY UNPINNED at the per-pixel level: see `oracle/__init__.py`.
"""

from __future__ import annotations

import struct
from pathlib import Path

import numpy as np
import torch
import torch.nn.functional as F

# ----------------------------------------------------------------------------------------------
# schema-free protobuf walk
# ----------------------------------------------------------------------------------------------


def _walk(data: bytes) -> dict[int, list]:
    out: dict[int, list] = {}
    i, n = 0, len(data)
    while i < n:
        key = 0
        shift = 0
        while True:
            b = data[i]
            i += 1
            key |= (b & 127) << shift
            shift += 7
            if b < 128:
                break
        fnum, wt = key >> 3, key & 7
        if wt == 0:
            v = 0
            shift = 0
            while True:
                b = data[i]
                i += 1
                v |= (b & 127) << shift
                shift += 7
                if b < 128:
                    break
        elif wt == 1:
            v = data[i : i + 8]
            i += 8
        elif wt == 2:
            ln = 0
            shift = 0
            while True:
                b = data[i]
                i += 1
                ln |= (b & 127) << shift
                shift += 7
                if b < 128:
                    break
            v = data[i : i + ln]
            i += ln
        elif wt == 5:
            v = data[i : i + 4]
            i += 4
        else:
            raise ValueError(f"wire type {wt}")
        out.setdefault(fnum, []).append(v)
    return out


def _ints(vals) -> list[int]:
    res = []
    for v in vals:
        if isinstance(v, int):
            res.append(v)
        else:  # packed
            j = 0
            while j < len(v):
                x = 0
                shift = 0
                while True:
                    b = v[j]
                    j += 1
                    x |= (b & 127) << shift
                    shift += 7
                    if b < 128:
                        break
                res.append(x)
    return [x - (1 << 64) if x >= (1 << 63) else x for x in res]


_NP = {1: "<f4", 2: "u1", 3: "i1", 6: "<i4", 7: "<i8", 9: "?", 10: "<f2", 11: "<f8"}


def _tensor(msg: dict) -> tuple[str, np.ndarray]:
    dims = _ints(msg.get(1, []))
    dt = msg.get(2, [1])[0]
    name = msg.get(8, [b""])[0].decode()
    if 9 in msg:
        arr = np.frombuffer(msg[9][0], dtype=_NP[dt]).copy()
    elif dt == 1:
        vals = []
        for v in msg.get(4, []):
            vals.extend(np.frombuffer(v, dtype="<f4").tolist())
        arr = np.asarray(vals, dtype=np.float32)
    elif dt == 7:
        arr = np.asarray(_ints(msg.get(7, [])), dtype=np.int64)
    elif dt == 11:
        vals = []
        for v in msg.get(10, []):
            vals.extend(np.frombuffer(v, dtype="<f8").tolist())
        arr = np.asarray(vals, dtype=np.float64)
    else:
        arr = np.asarray(_ints(msg.get(5, []))).astype(_NP[dt])
    return name, arr.reshape(dims)


def _attr(msg: dict):
    name = msg[1][0].decode()
    typ = msg.get(20, [0])[0]
    if typ == 1 or (typ == 0 and 2 in msg):
        return name, struct.unpack("<f", msg[2][0])[0]
    if typ == 2 or (typ == 0 and 3 in msg):
        return name, _ints(msg[3])[0]
    if typ == 3 or (typ == 0 and 4 in msg):
        return name, msg[4][0].decode()
    if typ == 4 or (typ == 0 and 5 in msg):
        return name, _tensor(_walk(msg[5][0]))[1]
    if typ == 6 or (typ == 0 and 7 in msg):
        vals = []
        for v in msg.get(7, []):
            vals.extend(np.frombuffer(v, dtype="<f4").tolist())
        return name, vals
    if typ == 7 or (typ == 0 and 8 in msg):
        return name, _ints(msg.get(8, []))
    return name, None


def _value_info(msg: dict):
    name = msg[1][0].decode()
    shape = []
    tt = _walk(_walk(msg[2][0])[1][0])
    for dim in _walk(tt[2][0]).get(1, []) if 2 in tt else []:
        d = _walk(dim)
        if 1 in d:
            shape.append(_ints(d[1])[0])
        elif 2 in d:
            shape.append(d[2][0].decode())
        else:
            shape.append(None)
    return name, shape


class RefGraph:
    """Decoded ONNX graph: nodes, initializers, I/O signatures."""

    def __init__(self, path: str | Path):
        model = _walk(Path(path).read_bytes())
        g = _walk(model[7][0])
        self.initializers = dict(_tensor(_walk(t)) for t in g.get(5, []))
        self.nodes = []
        for nb in g.get(1, []):
            n = _walk(nb)
            self.nodes.append(
                {
                    "op": n[4][0].decode(),
                    "in": [x.decode() for x in n.get(1, [])],
                    "out": [x.decode() for x in n.get(2, [])],
                    "attrs": dict(_attr(_walk(a)) for a in n.get(5, [])),
                }
            )
        self.inputs = [vi for vi in (_value_info(_walk(v)) for v in g.get(11, [])) if vi[0] not in self.initializers]
        self.outputs = [_value_info(_walk(v)) for v in g.get(12, [])]


# ----------------------------------------------------------------------------------------------
# operator table (ONNX opset 13 semantics on torch CPU)
# ----------------------------------------------------------------------------------------------


def _pads_to_torch(pads, x):
    """ONNX pads [b0,b1,...,e0,e1,...] over spatial dims; returns symmetric padding or pre-pads x."""
    nd = len(pads) // 2
    beg, end = pads[:nd], pads[nd:]
    if list(beg) == list(end):
        return x, tuple(beg)
    fpad = []
    for d in reversed(range(nd)):
        fpad += [beg[d], end[d]]
    return F.pad(x, fpad), tuple([0] * nd)


def _auto_pad(attrs, x, k, s):
    mode = attrs.get("auto_pad", "NOTSET")
    if mode in ("NOTSET", ""):
        return attrs.get("pads", [0] * (2 * len(k)))
    if mode == "VALID":
        return [0] * (2 * len(k))
    beg, end = [], []
    for d, (kk, ss) in enumerate(zip(k, s)):
        size = x.shape[2 + d]
        total = max((-(-size // ss) - 1) * ss + kk - size, 0)
        lo = total // 2 if mode == "SAME_UPPER" else total - total // 2
        beg.append(lo)
        end.append(total - lo)
    return beg + end


def _op_conv(x, w, b=None, **a):
    k = a.get("kernel_shape", list(w.shape[2:]))
    s = a.get("strides", [1] * len(k))
    x, pad = _pads_to_torch(_auto_pad(a, x, k, s), x)
    return F.conv2d(x, w, b, stride=tuple(s), padding=pad, dilation=tuple(a.get("dilations", [1, 1])), groups=a.get("group", 1))


def _op_convt(x, w, b=None, **a):
    k = a.get("kernel_shape", list(w.shape[2:]))
    s = a.get("strides", [1] * len(k))
    pads = a.get("pads", [0] * 4)
    assert pads[:2] == pads[2:], "asymmetric ConvTranspose pads not supported by the oracle"
    return F.conv_transpose2d(
        x, w, b, stride=tuple(s), padding=tuple(pads[:2]), output_padding=tuple(a.get("output_padding", [0, 0])),
        groups=a.get("group", 1), dilation=tuple(a.get("dilations", [1, 1])),
    )


def _op_pool(kind):
    def run(x, **a):
        k = a["kernel_shape"]
        s = a.get("strides", [1] * len(k))
        x, pad = _pads_to_torch(_auto_pad(a, x, k, s), x)
        ceil = bool(a.get("ceil_mode", 0))
        if kind == "max":
            return F.max_pool2d(x, tuple(k), tuple(s), pad, ceil_mode=ceil)
        return F.avg_pool2d(x, tuple(k), tuple(s), pad, ceil_mode=ceil, count_include_pad=bool(a.get("count_include_pad", 0)))

    return run


def _op_resize(x, roi=None, scales=None, sizes=None, **a):
    mode = a.get("mode", "nearest")
    ctm = a.get("coordinate_transformation_mode", "half_pixel")
    if sizes is not None and sizes.numel():
        size = [int(v) for v in sizes.tolist()[2:]]
        sf = None
    else:
        sc = [float(v) for v in scales.tolist()]
        assert sc[0] == 1.0 and sc[1] == 1.0, f"Resize over N/C not supported: {sc}"
        size = [int(np.floor(x.shape[2 + i] * sc[2 + i])) for i in range(2)]
        sf = sc[2:]
    if mode == "nearest":
        assert ctm == "asymmetric" and a.get("nearest_mode", "round_prefer_floor") == "floor" or (
            sf is not None and all(float(v).is_integer() for v in sf)
        ), f"nearest Resize variant not supported: {ctm}/{a.get('nearest_mode')}"
        return F.interpolate(x, size=size, mode="nearest")
    if mode == "linear":
        if ctm == "align_corners":
            return F.interpolate(x, size=size, mode="bilinear", align_corners=True)
        if ctm == "asymmetric":
            # ONNX Resize: x_original = x_resized / scale; weights from the fractional part, upper neighbour clamped
            out = x
            for dim, n_out in ((2, size[0]), (3, size[1])):
                n_in = out.shape[dim]
                pos = torch.arange(n_out, dtype=out.dtype) * (n_in / n_out)
                i0 = pos.floor().long().clamp(max=n_in - 1)
                i1 = (i0 + 1).clamp(max=n_in - 1)
                w1 = (pos - i0).view([-1 if d == dim else 1 for d in range(4)])
                lo, hi = out.index_select(dim, i0), out.index_select(dim, i1)
                out = lo + (hi - lo) * w1
            return out
        assert ctm in ("half_pixel", "pytorch_half_pixel"), f"linear Resize ctm={ctm} not supported"
        return F.interpolate(x, size=size, mode="bilinear", align_corners=False)
    raise NotImplementedError(f"Resize mode {mode}")


def _op_upsample(x, scales, **a):
    return _op_resize(x, None, scales, None, mode=a.get("mode", "nearest"), coordinate_transformation_mode="asymmetric", nearest_mode="floor")


def _op_bn(x, scale, bias, mean, var, **a):
    return F.batch_norm(x, mean, var, scale, bias, training=False, eps=a.get("epsilon", 1e-5))


def _op_pad(x, pads, value=None, **a):
    assert a.get("mode", "constant") == "constant"
    p = [int(v) for v in pads.tolist()]
    nd = len(p) // 2
    fpad = []
    for d in reversed(range(nd)):
        fpad += [p[d], p[nd + d]]
    return F.pad(x, fpad, value=float(value) if value is not None and value.numel() else 0.0)


def _op_clip(x, lo=None, hi=None, **a):
    lo_v = float(lo) if lo is not None and lo.numel() else a.get("min")
    hi_v = float(hi) if hi is not None and hi.numel() else a.get("max")
    return torch.clamp(x, min=lo_v, max=hi_v)


def _op_reshape(x, shape, **a):
    tgt = [int(v) for v in shape.tolist()]
    tgt = [x.shape[i] if v == 0 else v for i, v in enumerate(tgt)]
    return x.reshape(tgt)


def _op_squeeze(x, axes=None, **a):
    ax = [int(v) for v in axes.tolist()] if axes is not None else a.get("axes")
    if ax is None:
        return x.squeeze()
    for d in sorted((d % x.dim() for d in ax), reverse=True):
        x = x.squeeze(d)
    return x


def _op_unsqueeze(x, axes=None, **a):
    ax = [int(v) for v in axes.tolist()] if axes is not None else a.get("axes")
    for d in sorted(ax):
        x = x.unsqueeze(d)
    return x


_OPS = {
    "Conv": _op_conv,
    "ConvTranspose": _op_convt,
    "MaxPool": _op_pool("max"),
    "AveragePool": _op_pool("avg"),
    "GlobalAveragePool": lambda x, **a: x.mean(dim=(2, 3), keepdim=True),
    "Relu": lambda x, **a: torch.relu(x),
    "LeakyRelu": lambda x, **a: F.leaky_relu(x, a.get("alpha", 0.01)),
    "Sigmoid": lambda x, **a: torch.sigmoid(x),
    "Tanh": lambda x, **a: torch.tanh(x),
    "Add": lambda x, y, **a: x + y,
    "Sub": lambda x, y, **a: x - y,
    "Mul": lambda x, y, **a: x * y,
    "Div": lambda x, y, **a: x / y,
    "Concat": lambda *xs, **a: torch.cat(xs, dim=a["axis"]),
    "Transpose": lambda x, **a: x.permute(*a["perm"]).contiguous(),
    "Identity": lambda x, **a: x,
    "Resize": _op_resize,
    "Upsample": _op_upsample,
    "BatchNormalization": _op_bn,
    "Pad": _op_pad,
    "Clip": _op_clip,
    "Reshape": _op_reshape,
    "Squeeze": _op_squeeze,
    "Unsqueeze": _op_unsqueeze,
    "Cast": lambda x, **a: x.to(torch.float32) if int(a.get("to", 1)) == 1 else x,
}


class RefSession:
    """Minimal stand-in for the parts of `onnxruntime.InferenceSession` the reference touches."""

    def __init__(self, path: str | Path, dtype: torch.dtype = torch.float32, threads: int | None = None):
        self.graph = RefGraph(path)
        self.dtype = dtype
        if threads is not None:
            torch.set_num_threads(int(threads))
        self._consts = {}
        for k, v in self.graph.initializers.items():
            t = torch.from_numpy(np.ascontiguousarray(v))
            self._consts[k] = t.to(dtype) if t.is_floating_point() else t

    def input_signature(self):
        return list(self.graph.inputs)

    def output_signature(self):
        return list(self.graph.outputs)

    def run(self, feeds: dict[str, np.ndarray], want: list[str] | None = None) -> list[np.ndarray]:
        env = dict(self._consts)
        for name, arr in feeds.items():
            env[name] = torch.from_numpy(np.ascontiguousarray(arr)).to(self.dtype)
        with torch.no_grad():
            for node in self.graph.nodes:
                op = node["op"]
                if op == "Constant":
                    v = node["attrs"].get("value")
                    t = torch.from_numpy(np.ascontiguousarray(v))
                    env[node["out"][0]] = t.to(self.dtype) if t.is_floating_point() and t.numel() > 8 else t
                    continue
                if op not in _OPS:
                    raise NotImplementedError(f"oracle: ONNX operator '{op}' is not implemented")
                args = [env[n] if n else None for n in node["in"]]
                if op in ("Resize", "Upsample", "Pad", "Clip", "Reshape", "Squeeze", "Unsqueeze"):
                    pass
                else:
                    args = [a.to(self.dtype) if (a is not None and a.is_floating_point()) else a for a in args]
                env[node["out"][0]] = _OPS[op](*args, **node["attrs"])
        names = want or [o[0] for o in self.graph.outputs]
        return [env[n].to(torch.float32 if self.dtype == torch.float32 else self.dtype).numpy() for n in names]
