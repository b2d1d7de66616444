"""ONNX -> plan lowering (CPU): reader/writer round trip, contract, fusion, numerics of the lowered plan."""

from __future__ import annotations

import struct

import numpy as np
import pytest

from floodsr_b200 import graph as G
from floodsr_b200.h1 import H1_PARAM_COUNT, build_h1_model
from floodsr_b200.onnx_io import OnnxNode, load_onnx, save_onnx
from floodsr_b200.synth import synth_tile
from oracle import preprocessing_np as pp
from oracle.engine_ref import OracleEngine
from oracle.onnx_ref import RefGraph
from tests.plan_exec import run_plan


def test_onnx_roundtrip_and_two_readers_agree(h1_model_fp):
    m = load_onnx(h1_model_fp)
    assert m.initializer_param_count() == H1_PARAM_COUNT == 12_045_568  # infer_test_tiles.ipynb cell 9
    assert (m.ir_version, m.opset) == (7, 13)
    assert [(i.name, i.shape[1:]) for i in m.inputs] == [("depth_lr", [32, 32, 1]), ("dem_hr", [512, 512, 1])]
    assert [(o.name, o.shape[1:]) for o in m.outputs] == [("depth_hr_pred", [512, 512, 1])]
    ref = RefGraph(h1_model_fp)  # the oracle's independent decoder
    assert [n["op"] for n in ref.nodes] == [n.op_type for n in m.nodes]
    assert set(ref.initializers) == set(m.initializers)
    for k, v in m.initializers.items():
        assert np.array_equal(ref.initializers[k], v), k
    for a, b in zip(ref.nodes, m.nodes):
        assert a["in"] == b.inputs and a["out"] == b.outputs
        for key, val in b.attrs.items():
            got = a["attrs"][key]
            if isinstance(val, np.ndarray):
                assert np.array_equal(got, val)
            elif isinstance(val, float):
                assert got == pytest.approx(val)
            else:
                assert got == val, (b.name, key)


def test_contract_matches_reference_rules(h1_model_fp):
    c = G.resolve_contract(load_onnx(h1_model_fp))
    assert (c.depth_lr_hwc, c.dem_hr_hwc, c.output_hwc, c.scale) == ((32, 32, 1), (512, 512, 1), (512, 512, 1), 16)
    assert (c.depth_input_name, c.dem_input_name, c.output_name) == ("depth_lr", "dem_hr", "depth_hr_pred")
    m = load_onnx(h1_model_fp)
    m.inputs[0].name = "depth"
    with pytest.raises(AssertionError, match="model input 'depth_lr' not found"):
        G.resolve_contract(m)
    m = load_onnx(h1_model_fp)
    m.inputs[1].shape = ["n", 512, 512, 3]
    with pytest.raises(AssertionError, match="channels must be 1"):
        G.resolve_contract(m)
    m = load_onnx(h1_model_fp)
    m.inputs[1].shape = ["n", 500, 500, 1]
    with pytest.raises(AssertionError, match="must match output shape"):
        G.resolve_contract(m)


def test_h1_lowers_to_fused_plan(h1_model_fp):
    lm = G.lower_onnx(h1_model_fp)
    kinds = [op.kind for op in lm.ops]
    # 9 blocks x 3 convs, 4 max-pools + the 16x DEM average-pool, 4 upsamples, convT, fused head
    assert kinds.count(G.OP_CONV) == 27 and kinds.count(G.OP_POOL) == 5 and kinds.count(G.OP_UPSAMPLE) == 4
    assert kinds.count(G.OP_CONVT) == 1 and kinds[-1] == G.OP_HEAD and kinds.count(G.OP_ELTWISE) == 0
    assert sum(1 for op in lm.ops if op.res >= 0) == 9  # residual adds folded into the body's last conv
    assert lm.macs_per_tile() == 3_070_820_352  # 6.1416 GFLOP/tile
    hdr = struct.unpack("<8i", lm.plan_bytes[:32])
    assert hdr[0] == G.PLAN_MAGIC and hdr[2] == len(lm.tensors) and hdr[3] == len(lm.ops) and hdr[4:7] == (32, 512, 16)
    assert len(lm.plan_bytes) == 32 + 12 * len(lm.tensors) + 64 * len(lm.ops)
    assert lm.weights.dtype == np.float32 and lm.weights.size >= H1_PARAM_COUNT
    head = lm.ops[-1]
    assert head.k == 3 and head.cout == 32 and lm.tensors[head.src0] == (512, 512, 32) and lm.tensors[head.src1] == (512, 512, 1)


def test_lowered_plan_matches_oracle_interpreter(h1_model_fp):
    lm = G.lower_onnx(h1_model_fp)
    eng = OracleEngine(h1_model_fp)
    depth, dem = synth_tile(3)
    dn = pp.scale_depth_log1p(depth, 5.0)[None]
    en = pp.normalize_dem(dem)[0][None]
    want = eng.forward_norm(dn, en)
    got = run_plan(lm, dn, en)
    assert np.abs(got - want).max() < 5e-6
    assert want.std() > 0.02  # the random-init net produces a non-trivial field


def _check_against_oracle(fp, seed=5, tol=5e-6):
    lm = G.lower_onnx(fp)
    depth, dem = synth_tile(seed)
    dn = pp.scale_depth_log1p(depth, 5.0)[None]
    en = pp.normalize_dem(dem)[0][None]
    want = OracleEngine(fp).forward_norm(dn, en)
    assert np.abs(run_plan(lm, dn, en) - want).max() < tol
    return lm


def test_shortcut_conv_emitted_after_the_main_branch_is_not_read_before_it_ran(tmp_path):
    """conv1, conv2, conv_sc, Add(conv2, conv_sc): the residual must be folded into conv_sc (reading conv2), never into
    conv2 (which would read conv_sc before it exists)."""
    m = build_h1_model(seed=4)
    add_i = next(i for i, n in enumerate(m.nodes) if n.op_type == "Add")
    add = m.nodes[add_i]
    y_name, b_name = add.inputs  # y = block input after the projection, b = second body conv
    c = m.initializers[next(n for n in m.nodes if n.outputs == [b_name]).inputs[1]].shape[0]
    rng = np.random.default_rng(1)
    m.initializers["sc_W"] = (rng.standard_normal((c, c, 1, 1)) * 0.1).astype(np.float32)
    m.initializers["sc_B"] = (rng.standard_normal(c) * 0.05).astype(np.float32)
    sc = OnnxNode("Conv", [y_name, "sc_W", "sc_B"], ["sc_out"], name="shortcut",
                  attrs={"dilations": [1, 1], "group": 1, "kernel_shape": [1, 1], "pads": [0, 0, 0, 0], "strides": [1, 1]})
    m.nodes.insert(add_i, sc)  # after the main branch, right before the Add
    add.inputs[:] = [b_name, "sc_out"]
    fp = tmp_path / "shortcut.onnx"
    save_onnx(m, fp)
    lm = _check_against_oracle(fp)
    born = {op.dst: i for i, op in enumerate(lm.ops)}
    for i, op in enumerate(lm.ops):
        for src in (op.src0, op.src1, op.res):
            assert src < 2 or born[src] < i, (i, op.name, src)
    sc_op = next(op for op in lm.ops if op.name == "shortcut")
    assert sc_op.res >= 0 and sc_op.k == 1


def test_batchnorm_and_bias_add_fold_into_conv(tmp_path):
    m = build_h1_model(seed=2)
    # insert BatchNormalization + constant Add after the first conv (before its Relu)
    idx = next(i for i, n in enumerate(m.nodes) if n.op_type == "Conv")
    conv = m.nodes[idx]
    cout = m.initializers[conv.inputs[1]].shape[0]
    rng = np.random.default_rng(0)
    for nm, arr in [("bn_s", rng.uniform(0.5, 1.5, cout)), ("bn_b", rng.normal(0, 0.1, cout)), ("bn_m", rng.normal(0, 0.1, cout)),
                    ("bn_v", rng.uniform(0.5, 2.0, cout)), ("add_c", rng.normal(0, 0.1, (1, cout, 1, 1)))]:
        m.initializers[nm] = np.asarray(arr, np.float32)
    conv_out = conv.outputs[0]
    relu = next(n for n in m.nodes if n.inputs and n.inputs[0] == conv_out)
    bn = OnnxNode("BatchNormalization", [conv_out, "bn_s", "bn_b", "bn_m", "bn_v"], ["bn_out"], name="bn", attrs={"epsilon": 1e-3})
    add = OnnxNode("Add", ["bn_out", "add_c"], ["add_out"], name="biasadd")
    relu.inputs[0] = "add_out"
    m.nodes[idx + 1 : idx + 1] = [bn, add]
    fp = tmp_path / "bn.onnx"
    save_onnx(m, fp)
    lm = G.lower_onnx(fp)
    assert [op.kind for op in lm.ops].count(G.OP_CONV) == 27 and lm.ops[-1].kind == G.OP_HEAD
    depth, dem = synth_tile(5)
    dn = pp.scale_depth_log1p(depth, 5.0)[None]
    en = pp.normalize_dem(dem)[0][None]
    want = OracleEngine(fp).forward_norm(dn, en)
    assert np.abs(run_plan(lm, dn, en) - want).max() < 5e-6


def test_unsupported_operator_fails_loudly(tmp_path):
    m = build_h1_model(seed=3)
    idx = next(i for i, n in enumerate(m.nodes) if n.op_type == "Relu")
    m.nodes[idx].op_type = "Softplus"
    fp = tmp_path / "bad.onnx"
    save_onnx(m, fp)
    with pytest.raises(NotImplementedError, match="Softplus"):
        G.lower_onnx(fp)


def test_garbage_file_is_rejected(tmp_path):
    fp = tmp_path / "model_infer_dummy.onnx"
    fp.write_text("dummy onnx placeholder\n")  # the reference's 23-byte stub (tests/data/model_infer_dummy.onnx)
    with pytest.raises((ValueError, AssertionError, NotImplementedError)):
        G.lower_onnx(fp)


# ---------------------------------------------------------------------------------------------------------
# operator spellings a tf2onnx export of the real model may use (floodsr/models/ResUNet_16x_DEM.py:12-24 only describes
# the network in prose): every variant is lowered and executed against the oracle's interpreter of the same file
# ---------------------------------------------------------------------------------------------------------


@pytest.mark.parametrize("variant", ["pad_valid", "stride2", "stride2_tf", "bilinear_half_pixel", "bilinear_align_corners",
                                     "bilinear_asymmetric", "unfolded_bn", "clip_sigmoid", "reshape_cast", "everything"])
def test_export_spelling_variants_lower_and_match_the_oracle(tmp_path, variant):
    from tests import h1_variants as V

    m = build_h1_model(seed=6)
    n_elt = 0
    if variant == "pad_valid":
        V.explicit_pad_before_valid_conv(m)
    elif variant == "stride2":
        V.strided_conv_downsampling(m, level=1)
    elif variant == "stride2_tf":
        V.strided_conv_downsampling(m, level=0, tf_style=True)
    elif variant.startswith("bilinear_"):
        V.bilinear_resize(m, variant[len("bilinear_"):], which=(0, 3))
    elif variant == "unfolded_bn":
        V.unfolded_batchnorm(m)
    elif variant == "clip_sigmoid":
        V.clip_and_sigmoid_activations(m)
        n_elt = 2  # Clip(0, 1.5) and Sigmoid are element-wise ops; Clip(0, +inf) folds into its convolution as Relu
    elif variant == "reshape_cast":
        V.noop_reshape_and_cast(m)
    else:
        m = V.everything()
        n_elt = 2
    fp = tmp_path / f"{variant}.onnx"
    save_onnx(m, fp)
    lm = _check_against_oracle(fp, tol=1e-5)
    kinds = [op.kind for op in lm.ops]
    assert kinds[-1] == G.OP_HEAD and kinds.count(G.OP_ELTWISE) == n_elt
    if variant.startswith("stride2") or variant == "everything":
        picks = [op for op in lm.ops if op.kind == G.OP_POOL and op.mode == G.POOL_PICK]
        assert picks and all(op.k == 2 and op.aux in (0, 1) for op in picks)
        # the activation after a strided convolution folds into the convolution in front of the pick
        assert all(lm.ops[lm.ops.index(op) - 1].act == G.ACT_RELU for op in picks)
    if variant == "unfolded_bn":
        assert kinds.count(G.OP_CONV) == 27  # Mul / Sub / Div folded into the weights and bias


def test_lowering_rejects_what_it_cannot_express(tmp_path):
    from tests import h1_variants as V

    m = build_h1_model(seed=6)
    conv = [n for n in m.nodes if n.op_type == "Conv"][3]
    conv.attrs["pads"] = [0, 0, 0, 0]  # a VALID 3x3 convolution shrinks the map: not this network family
    fp = tmp_path / "valid.onnx"
    save_onnx(m, fp)
    with pytest.raises(NotImplementedError, match="not a 'same' convolution"):
        G.lower_onnx(fp)
    m = V.bilinear_resize(build_h1_model(seed=6), "tf_crop_and_resize", which=(1,))
    fp = tmp_path / "crop.onnx"
    save_onnx(m, fp)
    with pytest.raises(NotImplementedError, match="coordinate_transformation_mode"):
        G.lower_onnx(fp)
