"""Lower a `model_infer.onnx` graph to the B200 engine's fused layer plan.

This replaces what `onnxruntime.InferenceSession(model)` does for the reference
(`floodsr/engine/ort.py:54`): read the graph + initializers and make them executable.  The lowering walks
the ONNX nodes in order and emits the ops of `csrc/fsr_plan.h` over per-tile NHWC tensors:

* tf2onnx layout handling: graph I/O is NHWC, the interior NCHW; device tensors are always NHWC, so the
  NHWC<->NCHW `Transpose` nodes only flip a logical-layout tag.
* `Conv` (+ `BatchNormalization`, + bias `Add`, + residual `Add`, + `Relu`/`LeakyRelu`) fuse into one
  CONV op; a channel `Concat` feeding a conv becomes its second source (never materialised).
* `MaxPool`/`AveragePool` with kernel == stride -> POOL; nearest `Resize`/`Upsample` -> UPSAMPLE;
  `ConvTranspose` with kernel == stride -> CONVT.
* The trailing `conv kxk -> act -> conv 1x1 (1 channel, linear)` pair becomes the fused HEAD op.

Anything outside this set raises `NotImplementedError` naming the node: there is no CPU fallback.
"""

from __future__ import annotations

import struct
from dataclasses import dataclass, field
from pathlib import Path
from typing import Any

import numpy as np

from floodsr_b200.onnx_io import OnnxModel, OnnxNode, load_onnx

PLAN_MAGIC = 0x50525346
PLAN_VERSION = 2
OP_CONV, OP_POOL, OP_UPSAMPLE, OP_CONVT, OP_ELTWISE, OP_HEAD = 1, 2, 3, 4, 5, 6
ACT_NONE, ACT_RELU, ACT_LEAKY, ACT_CLIP, ACT_SIGMOID = 0, 1, 2, 3, 4  # CLIP / SIGMOID only on ELTWISE ops
POOL_MAX, POOL_AVG, POOL_PICK = 0, 1, 2  # PICK: element (aux, aux) of every k x k block (a strided convolution's sampling)
UP_NEAREST, UP_LINEAR_HALF_PIXEL, UP_LINEAR_ALIGN_CORNERS, UP_LINEAR_ASYMMETRIC = 0, 1, 2, 3
_OP_NAMES = {1: "CONV", 2: "POOL", 3: "UPSAMPLE", 4: "CONVT", 5: "ELTWISE", 6: "HEAD"}


@dataclass
class PlanOp:
    kind: int
    src0: int = -1
    src1: int = -1
    res: int = -1
    dst: int = -1
    k: int = 0
    mode: int = 0
    cout: int = 0
    act: int = ACT_NONE
    alpha: float = 0.0                # LEAKY slope / CLIP lower bound
    beta: float = 0.0                 # CLIP upper bound
    aux: int = 0                      # POOL_PICK: offset inside the block
    weight: np.ndarray | None = None  # HWIO
    bias: np.ndarray | None = None
    weight2: np.ndarray | None = None  # HEAD 1x1 [cmid]
    bias2: np.ndarray | None = None
    name: str = ""


@dataclass
class ModelContract:
    """Same fields as the reference's `ModelIOContract` (`floodsr/engine/ort.py:15-25`)."""

    depth_input_name: str
    dem_input_name: str
    output_name: str
    depth_lr_hwc: tuple[int, int, int]
    dem_hr_hwc: tuple[int, int, int]
    output_hwc: tuple[int, int, int]
    scale: int


@dataclass
class LoweredModel:
    contract: ModelContract
    tensors: list[tuple[int, int, int]]
    ops: list[PlanOp]
    out_tensor: int
    plan_bytes: bytes = b""
    weights: np.ndarray = field(default_factory=lambda: np.zeros(0, np.float32))

    def macs_per_tile(self) -> int:
        """Multiply-accumulates of all conv-like ops for one tile (roofline numerator / 2)."""
        total = 0
        for op in self.ops:
            if op.kind in (OP_CONV, OP_HEAD):
                h, w, _ = self.tensors[op.dst]
                cin = self.tensors[op.src0][2] + (self.tensors[op.src1][2] if op.src1 >= 0 else 0)
                total += h * w * op.k * op.k * cin * op.cout
                if op.kind == OP_HEAD:
                    total += h * w * op.cout
            elif op.kind == OP_CONVT:
                h, w, _ = self.tensors[op.dst]
                total += h * w * self.tensors[op.src0][2] * op.cout
        return int(total)

    def describe(self) -> list[str]:
        lines = []
        for i, op in enumerate(self.ops):
            srcs = [s for s in (op.src0, op.src1) if s >= 0]
            lines.append(
                f"{i:3d} {_OP_NAMES[op.kind]:8s} k={op.k} "
                + " + ".join(f"t{s}{self.tensors[s]}" for s in srcs)
                + (f" (+res t{op.res})" if op.res >= 0 else "")
                + f" -> t{op.dst}{self.tensors[op.dst]} act={op.act} {op.name}"
            )
        return lines


def _resolve_hwc(dims: list[Any], tensor_name: str) -> tuple[int, int, int]:
    """Same validation and wording as `EngineORT._resolve_hwc` (`ort.py:66-73`)."""
    assert len(dims) == 4, f"{tensor_name} must be rank-4 NHWC; got {dims}"
    h, w, c = dims[1], dims[2], dims[3]
    assert isinstance(h, int) and h > 0, f"{tensor_name} height must be fixed int; got {h}"
    assert isinstance(w, int) and w > 0, f"{tensor_name} width must be fixed int; got {w}"
    assert isinstance(c, int) and c == 1, f"{tensor_name} channels must be 1; got {c}"
    return (h, w, c)


def resolve_contract(model: OnnxModel) -> ModelContract:
    """Graph I/O contract, as `EngineORT._resolve_contract` derives it from ORT metadata (`ort.py:75-102`)."""
    input_meta = {vi.name: list(vi.shape) for vi in model.inputs}
    assert "depth_lr" in input_meta, "model input 'depth_lr' not found"
    assert "dem_hr" in input_meta, "model input 'dem_hr' not found"
    assert len(model.outputs) > 0, "model outputs are empty"
    for name in input_meta:
        if name not in ("depth_lr", "dem_hr"):
            raise AssertionError(f"unexpected model input name: {name}")
    out = model.outputs[0]
    depth_lr_hwc = _resolve_hwc(input_meta["depth_lr"], "depth_lr")
    dem_hr_hwc = _resolve_hwc(input_meta["dem_hr"], "dem_hr")
    output_hwc = _resolve_hwc(list(out.shape), out.name)
    assert dem_hr_hwc == output_hwc, f"DEM input shape {dem_hr_hwc} must match output shape {output_hwc}"
    assert dem_hr_hwc[0] % depth_lr_hwc[0] == 0, (
        f"HR/LR height ratio must be integer; got HR={dem_hr_hwc}, LR={depth_lr_hwc}"
    )
    return ModelContract(
        depth_input_name="depth_lr",
        dem_input_name="dem_hr",
        output_name=out.name,
        depth_lr_hwc=depth_lr_hwc,
        dem_hr_hwc=dem_hr_hwc,
        output_hwc=output_hwc,
        scale=int(dem_hr_hwc[0] // depth_lr_hwc[0]),
    )


@dataclass
class _Val:
    """An ONNX value during lowering: device tensor ids (several = pending channel concat) + logical layout."""

    tids: list[int]
    layout: str  # "NHWC" | "NCHW"
    producer: int = -1  # index into ops of the op that wrote tids[0] (single-tensor values only)
    pad: tuple[int, int, int, int] = (0, 0, 0, 0)  # pending zero padding (top, left, bottom, right) from a Pad node, consumed by
                                                   # the convolution that follows (tf2onnx emits SAME padding of strided convs so)


class _Lowering:
    def __init__(self, model: OnnxModel):
        self.m = model
        self.contract = resolve_contract(model)
        self.tensors: list[tuple[int, int, int]] = []
        self.ops: list[PlanOp] = []
        self.vals: dict[str, _Val] = {}
        self.consts: dict[str, np.ndarray] = dict(model.initializers)
        self.uses: dict[str, int] = {}
        for n in model.nodes:
            for i in n.inputs:
                self.uses[i] = self.uses.get(i, 0) + 1
        for o in model.outputs:
            self.uses[o.name] = self.uses.get(o.name, 0) + 1

    # -- helpers ---------------------------------------------------------------------------------
    def new_tensor(self, h: int, w: int, c: int) -> int:
        self.tensors.append((int(h), int(w), int(c)))
        return len(self.tensors) - 1

    def fail(self, node: OnnxNode, why: str):
        raise NotImplementedError(f"ONNX node '{node.name or node.outputs[0]}' ({node.op_type}): {why}")

    def act_in(self, node: OnnxNode, idx: int = 0) -> _Val:
        name = node.inputs[idx]
        if name not in self.vals:
            self.fail(node, f"input '{name}' is not an activation tensor")
        if any(self.vals[name].pad) and node.op_type != "Conv":
            self.fail(node, "a Pad node must be followed by the convolution it pads")
        return self.vals[name]

    def single(self, node: OnnxNode, v: _Val) -> int:
        if len(v.tids) != 1:
            self.fail(node, "channel Concat feeding this operator would have to be materialised")
        return v.tids[0]

    def fusable_producer(self, name: str, kinds=(OP_CONV, OP_CONVT, OP_ELTWISE), through_pick: bool = True) -> PlanOp | None:
        """The op that produced ONNX value `name`, if later nodes may still be folded into it.

        Pointwise followers (activation, per-channel scale / bias) commute with the element picking of a strided
        convolution, so they fold into the convolution in front of its POOL_PICK (`through_pick`)."""
        v = self.vals.get(name)
        if v is None or len(v.tids) != 1 or v.producer < 0 or self.uses.get(name, 0) != 1:
            return None
        op = self.ops[v.producer]
        if through_pick and op.kind == OP_POOL and op.mode == POOL_PICK and v.producer >= 1 and self.ops[v.producer - 1].dst == op.src0:
            op = self.ops[v.producer - 1]
        return op if op.kind in kinds else None

    # -- node handlers ---------------------------------------------------------------------------
    def lower(self) -> LoweredModel:
        c = self.contract
        t_depth = self.new_tensor(*c.depth_lr_hwc)
        t_dem = self.new_tensor(*c.dem_hr_hwc)
        self.vals["depth_lr"] = _Val([t_depth], "NHWC")
        self.vals["dem_hr"] = _Val([t_dem], "NHWC")
        for node in self.m.nodes:
            handler = getattr(self, "n_" + node.op_type, None)
            if handler is None:
                self.fail(node, "operator not supported by the B200 lowering")
            handler(node)
        out_name = self.m.outputs[0].name
        if out_name not in self.vals:
            raise NotImplementedError(f"graph output '{out_name}' was not produced by a supported operator")
        out = self.vals[out_name]
        if out.layout != "NHWC" or len(out.tids) != 1:
            raise NotImplementedError("graph output must be a single NHWC tensor")
        self.fuse_head(out.tids[0])
        out_t = out.tids[0]
        assert self.tensors[out_t] == c.output_hwc, f"lowered output shape {self.tensors[out_t]} != {c.output_hwc}"
        lowered = LoweredModel(contract=c, tensors=self.tensors, ops=self.ops, out_tensor=out_t)
        _serialise(lowered)
        return lowered

    def n_Constant(self, node: OnnxNode):
        v = node.attrs.get("value")
        if v is None:
            self.fail(node, "only tensor-valued Constant nodes are supported")
        self.consts[node.outputs[0]] = np.asarray(v)

    def n_Identity(self, node: OnnxNode):
        if node.inputs[0] in self.consts:
            self.consts[node.outputs[0]] = self.consts[node.inputs[0]]
        else:
            self.vals[node.outputs[0]] = self.act_in(node)

    def n_Transpose(self, node: OnnxNode):
        v = self.act_in(node)
        perm = list(node.attrs.get("perm", []))
        if perm == [0, 3, 1, 2] and v.layout == "NHWC":
            self.vals[node.outputs[0]] = _Val(v.tids, "NCHW", v.producer)
        elif perm == [0, 2, 3, 1] and v.layout == "NCHW":
            self.vals[node.outputs[0]] = _Val(v.tids, "NHWC", v.producer)
        else:
            self.fail(node, f"perm {perm} on a {v.layout} value is not a layout flip")

    def _check_nchw(self, node: OnnxNode, v: _Val):
        if v.layout != "NCHW":
            self.fail(node, "spatial operators are expected on NCHW-tagged values (tf2onnx interior)")

    def n_Conv(self, node: OnnxNode):
        v = self.act_in(node)
        self._check_nchw(node, v)
        if len(v.tids) > 2:
            self.fail(node, "conv over a concat of more than two tensors")
        w = self.consts.get(node.inputs[1])
        if w is None or w.ndim != 4:
            self.fail(node, "weights must be a rank-4 initializer")
        cout, cin, kh, kw = w.shape
        a = node.attrs
        strides = list(a.get("strides", [1, 1]))
        if a.get("group", 1) != 1 or kh != kw or kh not in (1, 3) or strides[0] != strides[1] or strides[0] not in (1, 2):
            self.fail(node, "only group=1, 1x1 or 3x3 convolutions with stride 1 or 2 are supported")
        if list(a.get("dilations", [1, 1])) != [1, 1]:
            self.fail(node, "dilated convolution")
        st = int(strides[0])
        h, wd, _ = self.tensors[v.tids[0]]
        # total padding = the node's own (pads / auto_pad) + what a preceding Pad node left pending on the value
        auto = a.get("auto_pad", "NOTSET")
        if auto in ("SAME_UPPER", "SAME_LOWER"):
            own = []
            for size in (h, wd):
                total = max((-(-size // st) - 1) * st + kh - size, 0)
                lo = total // 2 if auto == "SAME_UPPER" else total - total // 2
                own.append((lo, total - lo))
            pads = [own[0][0], own[1][0], own[0][1], own[1][1]]
        elif auto == "VALID":
            pads = [0, 0, 0, 0]
        else:
            pads = [int(x) for x in a.get("pads", [0, 0, 0, 0])]
        pt, pl, pb, pr = (pads[i] + v.pad[i] for i in range(4))
        # A stride-s convolution samples the stride-1 'same' convolution at rows pt' + s*i where pt' = k//2 - pt is how far the
        # first window's centre sits from row 0: lowered as CONV (stride 1, same padding) + POOL_PICK(k = s, offset pt').
        half = kh // 2
        off_y, off_x = half - pt, half - pl
        out_h = (h + pt + pb - kh) // st + 1
        out_w = (wd + pl + pr - kh) // st + 1
        ok = off_y == off_x and 0 <= off_y < st and out_h * st == h and out_w * st == wd
        if st == 1:
            ok = (pt, pl, pb, pr) == (half, half, half, half)
        if not ok:
            self.fail(node, f"padding {(pt, pl, pb, pr)} with stride {st} is not a 'same' convolution of the {h}x{wd} map")
        cin_have = sum(self.tensors[t][2] for t in v.tids)
        if cin_have != cin:
            self.fail(node, f"weight expects {cin} input channels, activation has {cin_have}")
        bias = None
        if len(node.inputs) > 2 and node.inputs[2]:
            bias = np.asarray(self.consts[node.inputs[2]], dtype=np.float32).reshape(cout)
        dst = self.new_tensor(h, wd, cout)
        op = PlanOp(
            OP_CONV, src0=v.tids[0], src1=v.tids[1] if len(v.tids) == 2 else -1, dst=dst, k=int(kh), cout=int(cout),
            weight=np.ascontiguousarray(np.transpose(np.asarray(w, np.float32), (2, 3, 1, 0))), bias=bias,
            name=node.name or node.outputs[0],
        )
        self.ops.append(op)
        if st == 1:
            self.vals[node.outputs[0]] = _Val([dst], "NCHW", len(self.ops) - 1)
            return
        # a value that is sampled afterwards can take no later fusion (activation, residual) before the sampling: those
        # commute with picking elements, so they are applied to the small tensor by ELTWISE ops if they follow
        small = self.new_tensor(h // st, wd // st, cout)
        self.ops.append(PlanOp(OP_POOL, src0=dst, dst=small, k=st, mode=POOL_PICK, aux=int(off_y), name=(node.name or node.outputs[0]) + "/stride"))
        self.vals[node.outputs[0]] = _Val([small], "NCHW", len(self.ops) - 1)

    def n_Pad(self, node: OnnxNode):
        v = self.act_in(node)
        self._check_nchw(node, v)
        pads = self.consts.get(node.inputs[1]) if len(node.inputs) > 1 and node.inputs[1] else node.attrs.get("pads")
        if pads is None:
            self.fail(node, "Pad needs constant pads")
        p = [int(x) for x in np.asarray(pads).reshape(-1)]
        value = 0.0
        if len(node.inputs) > 2 and node.inputs[2]:
            cv = self.consts.get(node.inputs[2])
            value = float(np.asarray(cv).reshape(-1)[0]) if cv is not None and np.asarray(cv).size else 0.0
        elif "value" in node.attrs:
            value = float(node.attrs["value"])
        if node.attrs.get("mode", "constant") != "constant" or value != 0.0 or len(p) != 8 or p[0] or p[1] or p[4] or p[5]:
            self.fail(node, "only zero-valued spatial Pad is supported")
        if any(x < 0 for x in p) or self.uses.get(node.outputs[0], 0) != 1:
            self.fail(node, "Pad must be non-negative and feed exactly one convolution")
        self.vals[node.outputs[0]] = _Val(v.tids, v.layout, -1, (v.pad[0] + p[2], v.pad[1] + p[3], v.pad[2] + p[6], v.pad[3] + p[7]))

    def n_ConvTranspose(self, node: OnnxNode):
        v = self.act_in(node)
        self._check_nchw(node, v)
        src = self.single(node, v)
        w = self.consts.get(node.inputs[1])
        if w is None or w.ndim != 4:
            self.fail(node, "weights must be a rank-4 initializer")
        cin, cout, kh, kw = w.shape
        a = node.attrs
        s = list(a.get("strides", [1, 1]))
        if a.get("group", 1) != 1 or kh != kw or s != [kh, kw] or any(a.get("pads", [0, 0, 0, 0])):
            self.fail(node, "only kernel == stride, unpadded transposed convolutions are supported")
        if any(a.get("output_padding", [0, 0])) or list(a.get("dilations", [1, 1])) != [1, 1]:
            self.fail(node, "output_padding / dilation")
        h, wd, c = self.tensors[src]
        if c != cin:
            self.fail(node, f"weight expects {cin} input channels, activation has {c}")
        bias = None
        if len(node.inputs) > 2 and node.inputs[2]:
            bias = np.asarray(self.consts[node.inputs[2]], dtype=np.float32).reshape(cout)
        dst = self.new_tensor(h * kh, wd * kw, cout)
        self.ops.append(
            PlanOp(OP_CONVT, src0=src, dst=dst, k=int(kh), cout=int(cout),
                   weight=np.ascontiguousarray(np.transpose(np.asarray(w, np.float32), (2, 3, 0, 1))), bias=bias,
                   name=node.name or node.outputs[0])
        )
        self.vals[node.outputs[0]] = _Val([dst], "NCHW", len(self.ops) - 1)

    def n_BatchNormalization(self, node: OnnxNode):
        prod = self.fusable_producer(node.inputs[0], kinds=(OP_CONV, OP_CONVT))
        if prod is None or prod.act != ACT_NONE or prod.res >= 0:
            self.fail(node, "BatchNormalization can only be folded into a directly preceding convolution")
        scale, bias, mean, var = (np.asarray(self.consts[n], np.float64) for n in node.inputs[1:5])
        g = scale / np.sqrt(var + float(node.attrs.get("epsilon", 1e-5)))
        prod.weight = (prod.weight.astype(np.float64) * g.reshape(1, 1, 1, -1)).astype(np.float32)
        b0 = prod.bias.astype(np.float64) if prod.bias is not None else np.zeros_like(g)
        prod.bias = ((b0 - mean) * g + bias).astype(np.float32)
        self.vals[node.outputs[0]] = self.vals[node.inputs[0]]

    def _activation(self, node: OnnxNode, act: int, alpha: float = 0.0):
        prod = self.fusable_producer(node.inputs[0])
        v = self.act_in(node)
        if prod is not None and prod.act == ACT_NONE:
            prod.act, prod.alpha = act, float(alpha)
            self.vals[node.outputs[0]] = v
            return
        src = self.single(node, v)
        dst = self.new_tensor(*self.tensors[src])
        self.ops.append(PlanOp(OP_ELTWISE, src0=src, dst=dst, act=act, alpha=float(alpha), name=node.name))
        self.vals[node.outputs[0]] = _Val([dst], v.layout, len(self.ops) - 1)

    def _pointwise(self, node: OnnxNode, act: int, alpha: float, beta: float):
        """Activations the convolution epilogues do not implement: a separate element-wise pass."""
        v = self.act_in(node)
        src = self.single(node, v)
        dst = self.new_tensor(*self.tensors[src])
        self.ops.append(PlanOp(OP_ELTWISE, src0=src, dst=dst, act=act, alpha=float(alpha), beta=float(beta), name=node.name))
        self.vals[node.outputs[0]] = _Val([dst], v.layout, len(self.ops) - 1)

    def n_Sigmoid(self, node: OnnxNode):
        self._pointwise(node, ACT_SIGMOID, 0.0, 0.0)

    def n_Clip(self, node: OnnxNode):
        def bound(i, key, default):
            if len(node.inputs) > i and node.inputs[i]:
                c = self.consts.get(node.inputs[i])
                if c is None:
                    self.fail(node, "Clip bounds must be constants")
                return float(np.asarray(c).reshape(-1)[0])
            return float(node.attrs.get(key, default))

        lo, hi = bound(1, "min", -3.4028234663852886e38), bound(2, "max", 3.4028234663852886e38)
        if lo == 0.0 and hi >= 3.0e38:
            self._activation(node, ACT_RELU)  # Clip(0, +inf) is how some exporters write Relu
        else:
            self._pointwise(node, ACT_CLIP, lo, hi)

    def _scale_shift(self, node: OnnxNode, kind: str):
        """Mul / Div / Sub by a per-channel (or scalar) constant right after a convolution: folded into its weights and bias
        (an unfolded BatchNormalization looks like this).  `const - x` and `const / x` are not affine in the weights."""
        a_name, b_name = node.inputs[0], node.inputs[1]
        if b_name in self.consts and a_name in self.vals:
            act_name, cst = a_name, self.consts[b_name]
        elif a_name in self.consts and b_name in self.vals and kind == "Mul":
            act_name, cst = b_name, self.consts[a_name]
        else:
            self.fail(node, f"{kind} needs (activation, constant) operands")
        prod = self.fusable_producer(act_name, kinds=(OP_CONV, OP_CONVT))
        c = np.asarray(cst, np.float64).reshape(-1)
        if prod is None or prod.act != ACT_NONE or prod.res >= 0 or c.size not in (1, prod.cout):
            self.fail(node, f"constant {kind} that is not a per-channel scale / shift right after a convolution")
        c = np.broadcast_to(c, (prod.cout,))
        if kind == "Sub":
            b0 = prod.bias.astype(np.float64) if prod.bias is not None else np.zeros(prod.cout)
            prod.bias = (b0 - c).astype(np.float32)
        else:
            g = c if kind == "Mul" else 1.0 / c
            prod.weight = (prod.weight.astype(np.float64) * g.reshape(1, 1, 1, -1)).astype(np.float32)
            if prod.bias is not None:
                prod.bias = (prod.bias.astype(np.float64) * g).astype(np.float32)
        self.vals[node.outputs[0]] = self.vals[act_name]

    def n_Mul(self, node: OnnxNode):
        self._scale_shift(node, "Mul")

    def n_Div(self, node: OnnxNode):
        self._scale_shift(node, "Div")

    def n_Sub(self, node: OnnxNode):
        self._scale_shift(node, "Sub")

    def _same_shape_view(self, node: OnnxNode, target: list[int] | None):
        """Reshape / Squeeze / Unsqueeze / Cast / Flatten-like nodes that leave a 4-D activation exactly as it is."""
        v = self.act_in(node)
        t = self.single(node, v)
        h, w, c = self.tensors[t]
        cur = [c, h, w] if v.layout == "NCHW" else [h, w, c]
        if target is not None:
            tgt = [int(x) for x in target]
            if len(tgt) != 4 or [cur[i] if x == 0 else x for i, x in enumerate(tgt[1:])] != cur or tgt[0] not in (0, -1, 1):
                self.fail(node, f"reshape of a {v.layout} {cur} activation to {tgt} changes its layout")
        self.vals[node.outputs[0]] = v

    def n_Reshape(self, node: OnnxNode):
        shape = self.consts.get(node.inputs[1]) if len(node.inputs) > 1 else None
        if shape is None:
            self.fail(node, "Reshape needs a constant shape")
        self._same_shape_view(node, list(np.asarray(shape).reshape(-1)))

    def n_Cast(self, node: OnnxNode):
        if int(node.attrs.get("to", 1)) != 1:
            self.fail(node, "only casts to float32 are supported")
        if node.inputs[0] in self.consts:
            self.consts[node.outputs[0]] = np.asarray(self.consts[node.inputs[0]], np.float32)
        else:
            self._same_shape_view(node, None)

    def n_Relu(self, node: OnnxNode):
        self._activation(node, ACT_RELU)

    def n_LeakyRelu(self, node: OnnxNode):
        self._activation(node, ACT_LEAKY, float(node.attrs.get("alpha", 0.01)))

    def n_Add(self, node: OnnxNode):
        a_name, b_name = node.inputs[0], node.inputs[1]
        # constant operand: per-channel bias folded into the producing convolution
        for act_name, const_name in ((a_name, b_name), (b_name, a_name)):
            if const_name in self.consts and act_name in self.vals:
                prod = self.fusable_producer(act_name, kinds=(OP_CONV, OP_CONVT))
                cst = np.asarray(self.consts[const_name], np.float32)
                if prod is None or prod.act != ACT_NONE or prod.res >= 0 or cst.size not in (1, prod.cout):
                    self.fail(node, "constant Add that is not a per-channel bias after a convolution")
                add = np.broadcast_to(cst.reshape(-1), (prod.cout,)).astype(np.float32)
                prod.bias = add.copy() if prod.bias is None else (prod.bias + add).astype(np.float32)
                self.vals[node.outputs[0]] = self.vals[act_name]
                return
        va, vb = self.act_in(node, 0), self.act_in(node, 1)
        ta, tb = self.single(node, va), self.single(node, vb)
        if self.tensors[ta] != self.tensors[tb] or va.layout != vb.layout:
            self.fail(node, "Add operands must have identical shapes (no broadcasting)")
        # residual add folded into the convolution that produced one operand -- only when the other operand already exists
        # at that point of the plan (a shortcut convolution emitted AFTER the main branch must not be read before it ran)
        def born(tid: int) -> int:
            return next((i for i, op in enumerate(self.ops) if op.dst == tid), -1)  # graph inputs: -1

        for mine, other, v in ((a_name, tb, va), (b_name, ta, vb)):
            prod = self.fusable_producer(mine, kinds=(OP_CONV,), through_pick=False)
            if prod is not None and prod.act == ACT_NONE and prod.res < 0 and born(other) < v.producer:
                prod.res = other
                self.vals[node.outputs[0]] = _Val(v.tids, v.layout, v.producer)
                return
        dst = self.new_tensor(*self.tensors[ta])
        self.ops.append(PlanOp(OP_ELTWISE, src0=ta, src1=tb, dst=dst, name=node.name))
        self.vals[node.outputs[0]] = _Val([dst], va.layout, len(self.ops) - 1)

    def n_Concat(self, node: OnnxNode):
        vs = [self.act_in(node, i) for i in range(len(node.inputs))]
        layout = vs[0].layout
        axis = int(node.attrs.get("axis", 1))
        chan_axis = 1 if layout == "NCHW" else 3
        if any(v.layout != layout for v in vs) or axis % 4 != chan_axis:
            self.fail(node, "only channel-axis Concat of same-layout tensors is supported")
        tids = [t for v in vs for t in v.tids]
        hw = {self.tensors[t][:2] for t in tids}
        if len(hw) != 1:
            self.fail(node, "Concat operands differ in spatial size")
        self.vals[node.outputs[0]] = _Val(tids, layout)

    def _pool(self, node: OnnxNode, mode: int):
        v = self.act_in(node)
        self._check_nchw(node, v)
        src = self.single(node, v)
        a = node.attrs
        k = list(a.get("kernel_shape", []))
        if len(k) != 2 or k[0] != k[1] or list(a.get("strides", [1, 1])) != k or any(a.get("pads", [0] * 4)):
            self.fail(node, "only square pooling with stride == kernel and no padding is supported")
        if a.get("ceil_mode", 0) or a.get("auto_pad", "NOTSET") not in ("NOTSET", "VALID", "SAME_UPPER"):
            self.fail(node, "ceil_mode / auto_pad pooling")
        h, w, c = self.tensors[src]
        if h % k[0] or w % k[0]:
            self.fail(node, "pooling window does not divide the input")
        dst = self.new_tensor(h // k[0], w // k[0], c)
        self.ops.append(PlanOp(OP_POOL, src0=src, dst=dst, k=int(k[0]), mode=mode, name=node.name))
        self.vals[node.outputs[0]] = _Val([dst], "NCHW", len(self.ops) - 1)

    def n_MaxPool(self, node: OnnxNode):
        self._pool(node, 0)

    def n_AveragePool(self, node: OnnxNode):
        self._pool(node, 1)

    def _upsample(self, node: OnnxNode, scales: np.ndarray | None, sizes: np.ndarray | None):
        v = self.act_in(node)
        self._check_nchw(node, v)
        src = self.single(node, v)
        h, w, c = self.tensors[src]
        mode = node.attrs.get("mode", "nearest")
        if isinstance(mode, bytes):
            mode = mode.decode()
        ctm = node.attrs.get("coordinate_transformation_mode", "half_pixel")
        if isinstance(ctm, bytes):
            ctm = ctm.decode()
        if mode == "nearest":
            up_mode = UP_NEAREST  # integer factors: asymmetric/floor and half_pixel/round_prefer_floor pick the same source pixel
        elif mode in ("linear", "bilinear"):
            up_mode = {"half_pixel": UP_LINEAR_HALF_PIXEL, "pytorch_half_pixel": UP_LINEAR_HALF_PIXEL,
                       "align_corners": UP_LINEAR_ALIGN_CORNERS, "asymmetric": UP_LINEAR_ASYMMETRIC}.get(ctm)
            if up_mode is None:
                self.fail(node, f"linear Resize with coordinate_transformation_mode={ctm}")
        else:
            self.fail(node, f"Resize mode '{mode}' is not supported")
        if scales is not None and scales.size == 4:
            sc = [float(x) for x in scales.reshape(-1)]
        elif sizes is not None and sizes.size == 4:
            sz = [int(x) for x in sizes.reshape(-1)]
            sc = [1.0, 1.0, sz[2] / h, sz[3] / w]
        else:
            self.fail(node, "Resize needs constant scales or sizes")
        if sc[0] != 1.0 or sc[1] != 1.0 or sc[2] != sc[3] or not float(sc[2]).is_integer() or sc[2] < 1:
            self.fail(node, f"only integer spatial upsampling is supported (scales={sc})")
        f = int(sc[2])
        dst = self.new_tensor(h * f, w * f, c)
        self.ops.append(PlanOp(OP_UPSAMPLE, src0=src, dst=dst, k=f, mode=up_mode, name=node.name))
        self.vals[node.outputs[0]] = _Val([dst], "NCHW", len(self.ops) - 1)

    def n_Resize(self, node: OnnxNode):
        def const(i):
            return self.consts.get(node.inputs[i]) if len(node.inputs) > i and node.inputs[i] else None

        self._upsample(node, const(2), const(3))

    def n_Upsample(self, node: OnnxNode):
        sc = self.consts.get(node.inputs[1]) if len(node.inputs) > 1 else np.asarray(node.attrs.get("scales", []))
        self._upsample(node, np.asarray(sc) if sc is not None else None, None)

    # -- head fusion -----------------------------------------------------------------------------
    def fuse_head(self, out_t: int):
        """conv kxk (2 sources, 1-channel second source) -> act -> conv 1x1 to one linear channel."""
        if len(self.ops) < 2:
            return
        last, prev = self.ops[-1], self.ops[-2]
        if not (last.kind == OP_CONV and last.k == 1 and last.cout == 1 and last.act == ACT_NONE and last.res < 0
                and last.src1 < 0 and last.dst == out_t and last.src0 == prev.dst):
            return
        if not (prev.kind == OP_CONV and prev.res < 0 and prev.src1 >= 0 and self.tensors[prev.src1][2] == 1):
            return
        users = sum(1 for op in self.ops if prev.dst in (op.src0, op.src1, op.res))
        if users != 1:
            return
        head = PlanOp(
            OP_HEAD, src0=prev.src0, src1=prev.src1, dst=out_t, k=prev.k, cout=prev.cout, act=prev.act, alpha=prev.alpha,
            weight=prev.weight, bias=prev.bias, weight2=last.weight.reshape(-1).astype(np.float32), bias2=last.bias,
            name=f"head({prev.name},{last.name})",
        )
        self.ops[-2:] = [head]


def _serialise(lm: LoweredModel) -> None:
    blobs: list[np.ndarray] = []
    off = 0

    def put(arr: np.ndarray | None) -> int:
        nonlocal off
        if arr is None:
            return -1
        a = np.ascontiguousarray(arr, dtype=np.float32).reshape(-1)
        # keep every blob 16-byte aligned for vectorised device loads
        pad = (-off) % 4
        if pad:
            blobs.append(np.zeros(pad, np.float32))
            off += pad
        start = off
        blobs.append(a)
        off += a.size
        return start

    c = lm.contract
    out = bytearray(
        struct.pack("<8i", PLAN_MAGIC, PLAN_VERSION, len(lm.tensors), len(lm.ops), c.depth_lr_hwc[0], c.dem_hr_hwc[0], c.scale, lm.out_tensor)
    )
    for h, w, ch in lm.tensors:
        out += struct.pack("<3i", h, w, ch)
    for op in lm.ops:
        w_off, b_off, w2_off, b2_off = put(op.weight), put(op.bias), put(op.weight2), put(op.bias2)
        out += struct.pack(
            "<9if4ifi", op.kind, op.src0, op.src1, op.res, op.dst, op.k, op.mode, op.cout, op.act, op.alpha,
            w_off, b_off, w2_off, b2_off, op.beta, op.aux,
        )
    lm.plan_bytes = bytes(out)
    lm.weights = np.concatenate(blobs) if blobs else np.zeros(0, np.float32)


def lower_onnx(model_or_path: OnnxModel | str | Path) -> LoweredModel:
    """Parse (if needed) and lower an ONNX model to the engine plan."""
    model = model_or_path if isinstance(model_or_path, OnnxModel) else load_onnx(model_or_path)
    return _Lowering(model).lower()
