"""Random-init `ResUNet_16x_DEM` graph written as a real ONNX file (offline stand-in for model_infer.onnx).

The reference never defines the network in code: it only ships the release asset `model_infer.onnx`
(`floodsr/models.json:4-6`) and describes it in prose (`floodsr/models/ResUNet_16x_DEM.py:5-24`):
dem_hr average-pooled 16x to LR and concatenated with depth_lr, a 4-level residual UNet with widths
f..16f, a 16x transposed convolution to HR, concat with dem_hr, 1-channel linear head.  The probe in
`proof_of_concepts/infer_test_tiles.ipynb` (cell 9) pins I/O names/shapes and the initializer element
count 12,045,568.  Hypothesis H1 is the one member of that family that hits the
count exactly; this module emits it in tf2onnx style (NHWC graph I/O, NCHW interior, OIHW conv weights,
bias as third Conv input) so that the engine's loader is exercised exactly as with the real asset.

Where the prose is silent (activation, pooling type, upsampling mode, residual wiring) H1 picks ReLU,
2x2 max-pool, nearest 2x Resize and `relu(x + body(x))`.  When the real asset is present nothing here is
used: the engine lowers whatever graph the file holds (`floodsr_b200/graph.py`).
"""

from __future__ import annotations

from pathlib import Path

import numpy as np

from floodsr_b200.onnx_io import DT_FLOAT, OnnxModel, OnnxNode, OnnxValueInfo, save_onnx

H1_PARAM_COUNT = 12_045_568
LR_TILE = 32
HR_TILE = 512
SCALE = 16


class _Builder:
    def __init__(self, rng: np.random.Generator):
        self.rng = rng
        self.nodes: list[OnnxNode] = []
        self.init: dict[str, np.ndarray] = {}
        self._n = 0

    def name(self, stem: str) -> str:
        self._n += 1
        return f"{stem}_{self._n}"

    def node(self, op: str, inputs: list[str], attrs: dict | None = None, stem: str | None = None) -> str:
        out = self.name(stem or op.lower())
        self.nodes.append(OnnxNode(op, list(inputs), [out], name=out + "_node", attrs=dict(attrs or {})))
        return out + ""

    def conv(self, x: str, cin: int, cout: int, k: int, *, bias: bool = True, gain: float = 1.0, stem: str = "conv") -> str:
        fan_in = cin * k * k
        w = self.rng.standard_normal((cout, cin, k, k), dtype=np.float32) * np.float32(gain * np.sqrt(2.0 / fan_in))
        wn = self.name(stem + "_W")
        self.init[wn] = w.astype(np.float32)
        ins = [x, wn]
        if bias:
            bn = self.name(stem + "_B")
            self.init[bn] = (self.rng.standard_normal(cout, dtype=np.float32) * np.float32(0.05)).astype(np.float32)
            ins.append(bn)
        pad = k // 2
        return self.node(
            "Conv",
            ins,
            {"dilations": [1, 1], "group": 1, "kernel_shape": [k, k], "pads": [pad, pad, pad, pad], "strides": [1, 1]},
            stem=stem,
        )

    def relu(self, x: str) -> str:
        return self.node("Relu", [x])

    def block(self, x: str, cin: int, c: int, stem: str) -> str:
        """Projection conv + 2-conv residual body: y = relu(conv(x)); out = relu(y + conv(relu(conv(y))))."""
        y = self.relu(self.conv(x, cin, c, 3, stem=stem + "_proj"))
        b = self.relu(self.conv(y, c, c, 3, stem=stem + "_b1"))
        b = self.conv(b, c, c, 3, gain=0.3, stem=stem + "_b2")
        return self.relu(self.node("Add", [y, b], stem=stem + "_add"))


def build_h1_model(seed: int = 0, base_filters: int = 32) -> OnnxModel:
    """Build the H1 graph with He-normal random weights (deterministic for a given seed)."""
    rng = np.random.default_rng(seed)
    b = _Builder(rng)
    f = base_filters

    # NHWC graph inputs -> NCHW interior, as tf2onnx does for Keras Conv2D models.
    depth = b.node("Transpose", ["depth_lr"], {"perm": [0, 3, 1, 2]}, stem="depth_nchw")
    dem = b.node("Transpose", ["dem_hr"], {"perm": [0, 3, 1, 2]}, stem="dem_nchw")
    dem_lr = b.node("AveragePool", [dem], {"kernel_shape": [SCALE, SCALE], "strides": [SCALE, SCALE]}, stem="dem_lr")
    x = b.node("Concat", [depth, dem_lr], {"axis": 1}, stem="enc_in")

    widths = [f, 2 * f, 4 * f, 8 * f, 16 * f]
    skips: list[tuple[str, int]] = []
    cin = 2
    for lvl, c in enumerate(widths):
        x = b.block(x, cin, c, f"enc{lvl}")
        cin = c
        if lvl < len(widths) - 1:
            skips.append((x, c))
            x = b.node("MaxPool", [x], {"kernel_shape": [2, 2], "strides": [2, 2]}, stem=f"pool{lvl}")

    # Resize operands as Constant nodes (not initializers) so the initializer count stays the probe's.
    b.nodes.append(OnnxNode("Constant", [], ["resize_roi"], name="resize_roi_c", attrs={"value": np.zeros((0,), dtype=np.float32)}))
    b.nodes.append(
        OnnxNode("Constant", [], ["resize_scales"], name="resize_scales_c", attrs={"value": np.asarray([1.0, 1.0, 2.0, 2.0], dtype=np.float32)})
    )
    for lvl in range(len(widths) - 2, -1, -1):
        skip, c = skips[lvl]
        x = b.node(
            "Resize",
            [x, "resize_roi", "resize_scales"],
            {"mode": "nearest", "coordinate_transformation_mode": "asymmetric", "nearest_mode": "floor"},
            stem=f"up{lvl}",
        )
        x = b.node("Concat", [x, skip], {"axis": 1}, stem=f"cat{lvl}")
        x = b.block(x, cin + c, c, f"dec{lvl}")
        cin = c

    # 16x transposed convolution (kernel == stride, no overlap between output patches).
    wt = rng.standard_normal((cin, f, SCALE, SCALE), dtype=np.float32) * np.float32(np.sqrt(2.0 / cin))
    b.init["up16_W"] = wt.astype(np.float32)
    b.init["up16_B"] = (rng.standard_normal(f, dtype=np.float32) * np.float32(0.05)).astype(np.float32)
    x = b.node(
        "ConvTranspose",
        [x, "up16_W", "up16_B"],
        {"dilations": [1, 1], "group": 1, "kernel_shape": [SCALE, SCALE], "pads": [0, 0, 0, 0], "strides": [SCALE, SCALE]},
        stem="up16",
    )
    x = b.relu(x)
    x = b.node("Concat", [x, dem], {"axis": 1}, stem="head_in")
    x = b.relu(b.conv(x, f + 1, f, 3, stem="head3"))
    # final linear 1x1 without bias; positive-mean weights keep the random net's output inside (0, 1)
    w1 = np.abs(rng.standard_normal((1, f, 1, 1), dtype=np.float32)) * np.float32(0.02)
    b.init["head1_W"] = w1.astype(np.float32)
    x = b.node(
        "Conv",
        [x, "head1_W"],
        {"dilations": [1, 1], "group": 1, "kernel_shape": [1, 1], "pads": [0, 0, 0, 0], "strides": [1, 1]},
        stem="head1",
    )
    b.nodes.append(OnnxNode("Transpose", [x], ["depth_hr_pred"], name="out_nhwc", attrs={"perm": [0, 2, 3, 1]}))

    return OnnxModel(
        nodes=b.nodes,
        initializers=b.init,
        inputs=[
            OnnxValueInfo("depth_lr", DT_FLOAT, ["unk__300", LR_TILE, LR_TILE, 1]),
            OnnxValueInfo("dem_hr", DT_FLOAT, ["unk__301", HR_TILE, HR_TILE, 1]),
        ],
        outputs=[OnnxValueInfo("depth_hr_pred", DT_FLOAT, ["unk__302", HR_TILE, HR_TILE, 1])],
        ir_version=7,
        opset=13,
        producer_name="floodsr_b200.h1",
        producer_version="1",
        graph_name="ResUNet_16x_DEM_h1",
    )


def write_h1_model(path: str | Path, seed: int = 0, output_gain: float | None = None) -> Path:
    """Write the random-init H1 network to `path` and return it (`output_gain` rescales the final 1x1 weights)."""
    model = build_h1_model(seed=seed)
    if output_gain is not None:
        model.initializers["head1_W"] = (model.initializers["head1_W"] * np.float32(output_gain)).astype(np.float32)
    path = Path(path)
    path.parent.mkdir(parents=True, exist_ok=True)
    save_onnx(model, path)
    return path
