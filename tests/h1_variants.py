"""Test helper: H1 rewritten with the operator spellings a tf2onnx export of the real model may use instead of H1's.

Every variant computes a DIFFERENT function from H1 (new random weights where a node is replaced), so each is checked
against the oracle's interpreter of the same file, never against H1's outputs.
"""

from __future__ import annotations

import numpy as np

from floodsr_b200.h1 import build_h1_model
from floodsr_b200.onnx_io import OnnxModel, OnnxNode

_CONV_ATTRS = {"dilations": [1, 1], "group": 1}


def _consumers(m: OnnxModel, name: str):
    return [n for n in m.nodes if name in n.inputs]


def _rewire(m: OnnxModel, old: str, new: str, skip=()):
    for n in m.nodes:
        if n in skip:
            continue
        n.inputs[:] = [new if i == old else i for i in n.inputs]


def explicit_pad_before_valid_conv(m: OnnxModel, which: int = 3) -> OnnxModel:
    """'same' 3x3 conv -> Pad(1) + VALID conv (how tf2onnx writes some SAME convolutions)."""
    conv = [n for n in m.nodes if n.op_type == "Conv" and n.attrs["kernel_shape"] == [3, 3]][which]
    m.initializers["pad1_pads"] = np.asarray([0, 0, 1, 1, 0, 0, 1, 1], np.int64)
    pad = OnnxNode("Pad", [conv.inputs[0], "pad1_pads"], ["pad1_out"], name="pad1", attrs={"mode": "constant"})
    conv.inputs[0] = "pad1_out"
    conv.attrs["pads"] = [0, 0, 0, 0]
    m.nodes.insert(m.nodes.index(conv), pad)
    return m


def strided_conv_downsampling(m: OnnxModel, level: int = 1, tf_style: bool = False, seed: int = 9) -> OnnxModel:
    """MaxPool of encoder `level` -> 3x3 stride-2 convolution (pads 1 all round, or tf SAME: Pad(0,0,1,1) + VALID)."""
    pool = [n for n in m.nodes if n.op_type == "MaxPool"][level]
    src = pool.inputs[0]
    c = next(m.initializers[p.inputs[1]].shape[0] for p in m.nodes if p.op_type == "Conv" and _feeds(m, p.outputs[0], src))
    rng = np.random.default_rng(seed)
    m.initializers[f"down{level}_W"] = (rng.standard_normal((c, c, 3, 3)) * np.sqrt(2.0 / (9 * c))).astype(np.float32)
    m.initializers[f"down{level}_B"] = (rng.standard_normal(c) * 0.05).astype(np.float32)
    i = m.nodes.index(pool)
    new = []
    x = src
    pads = [1, 1, 1, 1]
    if tf_style:
        m.initializers[f"down{level}_pads"] = np.asarray([0, 0, 0, 0, 0, 0, 1, 1], np.int64)
        new.append(OnnxNode("Pad", [src, f"down{level}_pads"], [f"down{level}_pad"], name=f"down{level}_padnode", attrs={"mode": "constant"}))
        x, pads = f"down{level}_pad", [0, 0, 0, 0]
    new.append(OnnxNode("Conv", [x, f"down{level}_W", f"down{level}_B"], [f"down{level}_conv"], name=f"down{level}",
                        attrs=dict(_CONV_ATTRS, kernel_shape=[3, 3], pads=pads, strides=[2, 2])))
    new.append(OnnxNode("Relu", [f"down{level}_conv"], pool.outputs[:], name=f"down{level}_relu"))
    m.nodes[i : i + 1] = new
    return m


def _feeds(m: OnnxModel, start: str, target: str, depth: int = 4) -> bool:
    """True when value `start` reaches `target` through at most `depth` single-input nodes (Relu / Add chains)."""
    frontier = {start}
    for _ in range(depth):
        if target in frontier:
            return True
        frontier = {o for n in m.nodes for o in n.outputs if any(i in frontier for i in n.inputs)}
    return target in frontier


def bilinear_resize(m: OnnxModel, ctm: str = "half_pixel", which=(0, 2)) -> OnnxModel:
    """Decoder Resize nodes -> mode linear with the given coordinate_transformation_mode."""
    ups = [n for n in m.nodes if n.op_type == "Resize"]
    for k in which:
        ups[k].attrs.update({"mode": "linear", "coordinate_transformation_mode": ctm})
        ups[k].attrs.pop("nearest_mode", None)
    return m


def unfolded_batchnorm(m: OnnxModel, which: int = 5, seed: int = 4) -> OnnxModel:
    """conv -> Mul(scale) -> Sub(mean') -> Div(c) -> Relu, the pieces an unfused normalisation leaves behind."""
    conv = [n for n in m.nodes if n.op_type == "Conv" and n.attrs["kernel_shape"] == [3, 3]][which]
    cout = m.initializers[conv.inputs[1]].shape[0]
    rng = np.random.default_rng(seed)
    m.initializers["bnu_scale"] = rng.uniform(0.6, 1.4, (1, cout, 1, 1)).astype(np.float32)
    m.initializers["bnu_shift"] = rng.normal(0, 0.1, (1, cout, 1, 1)).astype(np.float32)
    m.initializers["bnu_div"] = np.asarray(1.25, np.float32)
    out = conv.outputs[0]
    follow = _consumers(m, out)
    mul = OnnxNode("Mul", ["bnu_scale", out], ["bnu_mul"], name="bnu_mul_node")
    sub = OnnxNode("Sub", ["bnu_mul", "bnu_shift"], ["bnu_sub"], name="bnu_sub_node")
    div = OnnxNode("Div", ["bnu_sub", "bnu_div"], ["bnu_div_out"], name="bnu_div_node")
    for n in follow:
        n.inputs[:] = ["bnu_div_out" if i == out else i for i in n.inputs]
    i = m.nodes.index(conv)
    m.nodes[i + 1 : i + 1] = [mul, sub, div]
    return m


def clip_and_sigmoid_activations(m: OnnxModel) -> OnnxModel:
    """One Relu -> Clip(0, 6) (ReLU6, bounds as inputs), one Relu -> Clip(0, +inf) via attributes-free inputs, one -> Sigmoid."""
    relus = [n for n in m.nodes if n.op_type == "Relu"]
    m.initializers["clip_lo"] = np.asarray(0.0, np.float32)
    m.initializers["clip_hi"] = np.asarray(1.5, np.float32)  # low enough to actually clip the random net's activations
    m.initializers["clip_inf"] = np.asarray(3.4028234663852886e38, np.float32)
    a, b, c = relus[2], relus[7], relus[12]
    a.op_type, a.inputs = "Clip", [a.inputs[0], "clip_lo", "clip_hi"]
    b.op_type, b.inputs = "Clip", [b.inputs[0], "clip_lo", "clip_inf"]
    c.op_type = "Sigmoid"
    return m


def noop_reshape_and_cast(m: OnnxModel) -> OnnxModel:
    """A Reshape to the same NCHW shape (0 = keep) and a Cast to float in the middle of the encoder."""
    pool = [n for n in m.nodes if n.op_type == "MaxPool"][0]
    src = pool.inputs[0]
    m.initializers["rs_shape"] = np.asarray([0, 0, 32, 32], np.int64)
    rs = OnnxNode("Reshape", [src, "rs_shape"], ["rs_out"], name="rs_node")
    cast = OnnxNode("Cast", ["rs_out"], ["cast_out"], name="cast_node", attrs={"to": 1})
    pool.inputs[0] = "cast_out"
    i = m.nodes.index(pool)
    m.nodes[i:i] = [rs, cast]
    return m


def everything(seed: int = 6) -> OnnxModel:
    m = build_h1_model(seed=seed)
    explicit_pad_before_valid_conv(m)
    strided_conv_downsampling(m, level=1)
    strided_conv_downsampling(m, level=2, tf_style=True, seed=10)
    bilinear_resize(m, "half_pixel", which=(0,))
    bilinear_resize(m, "align_corners", which=(2,))
    unfolded_batchnorm(m)
    clip_and_sigmoid_activations(m)
    noop_reshape_and_cast(m)
    return m
