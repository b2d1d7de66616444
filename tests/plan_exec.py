"""Test helper: execute a lowered plan (floodsr_b200.graph.LoweredModel) with torch CPU ops.

Only used to check the ONNX -> plan lowering without a GPU; it deliberately mirrors what the CUDA engine
does op by op (NHWC tensors, HWIO weights, fused bias/residual/activation, fused head).
"""

from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from floodsr_b200 import graph as G


def _act(x, op):
    if op.act == G.ACT_RELU:
        return torch.relu(x)
    if op.act == G.ACT_LEAKY:
        return F.leaky_relu(x, op.alpha)
    if op.act == G.ACT_CLIP:
        return torch.clamp(x, op.alpha, op.beta)
    if op.act == G.ACT_SIGMOID:
        return torch.sigmoid(x)
    return x


def _conv(x_nhwc, w_hwio, bias, k):
    x = x_nhwc.permute(0, 3, 1, 2)
    w = torch.from_numpy(np.ascontiguousarray(w_hwio)).permute(3, 2, 0, 1).to(x.dtype)
    b = torch.from_numpy(bias).to(x.dtype) if bias is not None else None
    return F.conv2d(x, w, b, padding=k // 2).permute(0, 2, 3, 1)


def _bilinear(x_nhwc, f, mode):
    """The engine's bilinear upsampling (fsr_common.cuh: up_linear_coord), written out with explicit gathers."""
    n, h, w, c = x_nhwc.shape

    def coords(n_in):
        o = torch.arange(n_in * f, dtype=x_nhwc.dtype)
        if mode == G.UP_LINEAR_ALIGN_CORNERS:
            x = o * (n_in - 1) / (n_in * f - 1) if n_in * f > 1 else o * 0
        elif mode == G.UP_LINEAR_ASYMMETRIC:
            x = o / f
        else:
            x = (o + 0.5) / f - 0.5
        x = x.clamp(0, n_in - 1)
        i0 = x.floor().long()
        i1 = (i0 + 1).clamp(max=n_in - 1)
        return i0, i1, (x - i0)

    y0, y1, wy = coords(h)
    x0, x1, wx = coords(w)
    wy, wx = wy.view(1, -1, 1, 1), wx.view(1, 1, -1, 1)
    rows0, rows1 = x_nhwc[:, y0], x_nhwc[:, y1]
    top = rows0[:, :, x0] + (rows0[:, :, x1] - rows0[:, :, x0]) * wx
    bot = rows1[:, :, x0] + (rows1[:, :, x1] - rows1[:, :, x0]) * wx
    return top + (bot - top) * wy


def run_plan(lm: G.LoweredModel, depth_norm: np.ndarray, dem_norm: np.ndarray, dtype=torch.float32, return_all: bool = False):
    """depth_norm [B,lr,lr], dem_norm [B,hr,hr] -> pred_norm [B,hr,hr] (or every plan tensor, NHWC, with `return_all`).

    `dtype=torch.float64` gives the error-budget reference: the same fp32 inputs and weights, arithmetic in double."""
    t: dict[int, torch.Tensor] = {
        0: torch.from_numpy(np.asarray(depth_norm, np.float32))[..., None].to(dtype),
        1: torch.from_numpy(np.asarray(dem_norm, np.float32))[..., None].to(dtype),
    }
    with torch.no_grad():
        for op in lm.ops:
            if op.kind in (G.OP_CONV, G.OP_HEAD):
                x = t[op.src0] if op.src1 < 0 else torch.cat([t[op.src0], t[op.src1]], dim=3)
                y = _conv(x, op.weight, op.bias, op.k)
                if op.res >= 0:
                    y = y + t[op.res]
                y = _act(y, op)
                if op.kind == G.OP_HEAD:
                    y = (y * torch.from_numpy(op.weight2).to(dtype)).sum(dim=3, keepdim=True)
                    if op.bias2 is not None:
                        y = y + float(op.bias2.reshape(-1)[0])
            elif op.kind == G.OP_POOL:
                x = t[op.src0].permute(0, 3, 1, 2)
                if op.mode == G.POOL_PICK:
                    y = x[:, :, op.aux :: op.k, op.aux :: op.k].permute(0, 2, 3, 1)
                else:
                    y = (F.max_pool2d(x, op.k) if op.mode == 0 else F.avg_pool2d(x, op.k)).permute(0, 2, 3, 1)
            elif op.kind == G.OP_UPSAMPLE:
                if op.mode == G.UP_NEAREST:
                    y = t[op.src0].repeat_interleave(op.k, dim=1).repeat_interleave(op.k, dim=2)
                else:
                    y = _bilinear(t[op.src0], op.k, op.mode)
            elif op.kind == G.OP_CONVT:
                x = t[op.src0].permute(0, 3, 1, 2)
                w = torch.from_numpy(np.ascontiguousarray(op.weight)).permute(2, 3, 0, 1).to(dtype)  # [kh,kw,ci,co] -> [ci,co,kh,kw]
                b = torch.from_numpy(op.bias).to(dtype) if op.bias is not None else None
                y = _act(F.conv_transpose2d(x, w, b, stride=op.k).permute(0, 2, 3, 1), op)
            elif op.kind == G.OP_ELTWISE:
                y = t[op.src0] if op.src1 < 0 else t[op.src0] + t[op.src1]
                y = _act(y, op)
            else:
                raise AssertionError(op.kind)
            assert tuple(y.shape[1:]) == lm.tensors[op.dst], (op.name, y.shape, lm.tensors[op.dst])
            t[op.dst] = y.contiguous()
    if return_all:
        return {k: v.numpy() for k, v in t.items()}
    return t[lm.out_tensor][..., 0].numpy()
