"""The oracle itself: op-level known answers, structure pinned to the shipped graph, golden vectors."""
import hashlib
import os

import numpy as np
import pytest

from conftest import REFERENCE_DIR, needs_reference
from oracle import tf_ops as ops
from oracle.roomnet_oracle import RoomNetOracle, flat_len, spatial_trace, synthetic_suite


def test_relu6_and_avgpool_small():
    x = np.array([-1.0, 0.5, 7.0], np.float32)
    assert ops.relu6(x).tolist() == [0.0, 0.5, 6.0]
    a = np.arange(16, dtype=np.float32).reshape(1, 4, 4, 1)
    np.testing.assert_allclose(ops.avg_pool_valid(a, 3, 1)[0, :, :, 0], [[5, 6], [9, 10]])
    np.testing.assert_allclose(ops.avg_pool_valid(a, 4, 2)[0, :, :, 0], [[7.5]])
    np.testing.assert_allclose(ops.avg_pool_valid(a, 2, 2)[0, :, :, 0], [[2.5, 4.5], [10.5, 12.5]])


def test_conv_valid_is_cross_correlation():
    x = np.zeros((1, 4, 4, 1), np.float32)
    x[0, 1, 2, 0] = 1.0
    w = np.arange(9, dtype=np.float32).reshape(3, 3, 1, 1)
    y = ops.conv2d_valid(x, w)[0, :, :, 0]
    # y[i,j] = w[1-i, 2-j] (no kernel flip)
    np.testing.assert_allclose(y, [[w[1, 2, 0, 0], w[1, 1, 0, 0]], [w[0, 2, 0, 0], w[0, 1, 0, 0]]])
    rng = np.random.default_rng(0)
    x = rng.normal(size=(2, 9, 9, 5)).astype(np.float32)
    w = rng.normal(size=(3, 3, 5, 7)).astype(np.float32)
    np.testing.assert_allclose(ops.conv2d_valid(x, w, "torch"), ops.conv2d_valid(x, w), rtol=1e-4, atol=1e-4)


def test_legacy_bilinear_hand_computed():
    # 1-D case, in=5 -> out=2: scale 2.5, src = [0, 2.5] -> lo=[0,2], t=[0,0.5]
    x = np.array([0, 10, 20, 30, 40], np.float32).reshape(1, 1, 5, 1)
    y = ops.resize_bilinear_legacy(np.repeat(x, 5, axis=1), 2, 2)[0, :, :, 0]
    np.testing.assert_allclose(y, [[0, 25], [0, 25]])
    # in=3 -> out=4 (upsample): src = 0, .75, 1.5, 2.25 ; hi clamps at in-1
    x = np.array([0, 4, 8], np.float32).reshape(1, 1, 3, 1)
    y = ops.resize_bilinear_legacy(x, 1, 4)[0, 0, :, 0]
    np.testing.assert_allclose(y, [0, 3, 6, 8])
    # identity when sizes match; commutes with a per-channel affine (used by the BN fold)
    r = np.random.default_rng(1).normal(size=(1, 21, 21, 3)).astype(np.float64)
    np.testing.assert_array_equal(ops.resize_bilinear_legacy(r, 21, 21), r)
    a, b = np.array([2.0, -3.0, 0.5]), np.array([1.0, 0.0, -7.0])
    np.testing.assert_allclose(ops.resize_bilinear_legacy(r * a + b, 2, 2), ops.resize_bilinear_legacy(r, 2, 2) * a + b,
                               atol=1e-12)


def test_batchnorm_epsilon():
    x = np.ones((1, 1, 1, 2), np.float32)
    y = ops.batch_norm_inference(x, np.array([2.0, 1.0], np.float32), np.array([0.5, 0.0], np.float32),
                                 np.array([0.0, 1.0], np.float32), np.array([1.0, 1e-36], np.float32))
    np.testing.assert_allclose(y[0, 0, 0], [2 / np.sqrt(1.001) + 0.5, 0.0], rtol=1e-6)


def test_softmax_argmax_ties():
    l = np.array([[0, 0.6, 2.7, 0, 0, 0], [3.0, 3.0, 1.0, 0, 0, 0]], np.float32)
    sm = ops.softmax(l)
    assert ops.argmax_first(sm).tolist() == [2, 0]
    assert sm[0, 0] == sm[0, 3] == sm[0, 4] == sm[0, 5]


def test_spatial_trace_matches_survey_appendix_a():
    assert [(t["conv"], t["out"]) for t in spatial_trace(224)] == [
        (222, 220), (218, 215), (213, 210), (208, 205), (203, 100), (98, 48), (46, 46), (44, 21), (19, 8), (6, 2)]
    assert (flat_len(224), flat_len(300), flat_len(600)) == (64, 256, 3136)


def test_center_crop_quirk():
    c = RoomNetOracle.center_crop
    a = np.arange(480 * 641 * 3, dtype=np.uint32).reshape(480, 641, 3)
    assert c(a).shape == (480, 480, 3) and c(a)[0, 0, 0] == a[0, 80, 0]  # (641-480)//2 = 80
    b = np.arange(641 * 480 * 3, dtype=np.uint32).reshape(641, 480, 3)
    # (480-641)//2 = -81 (floor) -> abs = 81, not 80: the reference's rounding quirk for h > w
    assert c(b).shape == (480, 480, 3) and c(b)[0, 0, 0] == b[81, 0, 0]


def test_appendix_f_vectors(weights):
    o = RoomNetOracle(dtype=np.float64, weights=weights)
    imgs = np.stack([np.zeros((224, 224, 3), np.uint8), np.full((224, 224, 3), 255, np.uint8),
                     np.full((224, 224, 3), 128, np.uint8)])
    r = o.forward(o.normalise(imgs))
    want = [[-0.501158, 0.599534, 2.711391, -0.480079, -1.119013, -0.533861],
            [1.113234, 0.619842, -0.366923, 0.465560, -1.152357, 0.087116],
            [-0.882862, -0.352664, 3.101252, -1.025706, -0.940316, 0.509617]]
    np.testing.assert_allclose(r["pre_relu6"], want, atol=2e-6)
    assert r["argmax"].tolist() == [2, 0, 2]


def test_suite_is_stable(suite64, golden):
    md5 = [hashlib.md5(im.tobytes()).hexdigest() for im in suite64]
    assert md5 == list(golden["image_md5"]), "synthetic suite changed (cv2 version?) — regenerate tests/golden"


def test_fp32_oracle_matches_golden(weights, suite64, golden):
    o = RoomNetOracle(dtype=np.float32, weights=weights, conv_backend="torch")
    sel = np.arange(0, 64, 3)
    r = o.forward(o.normalise(suite64[sel]))
    assert np.array_equal(r["argmax"], golden["argmax"][sel])
    assert np.abs(r["logits"] - golden["logits"][sel]).max() < 2e-4
    assert np.abs(r["softmax"] - golden["softmax"][sel]).max() < 1e-4


def test_folded_oracle_is_exact(weights, suite64, golden):
    from oracle.fold import fold, folded_forward
    r = folded_forward(fold(weights), suite64[:6], dtype=np.float64, conv_backend="torch")
    assert np.abs(r["pre_relu6"] - golden["pre_relu6"][:6]).max() < 1e-10


@needs_reference
def test_restatement_equals_shipped_graph(weights, suite64):
    """Executes final_model/roomnet.meta node by node: the hand-written structure must be bit-identical."""
    from oracle.tf_graph_interp import run_reference_graph
    o = RoomNetOracle(dtype=np.float32, weights=weights)
    x = o.normalise(suite64[:3])
    out, interp = run_reference_graph(x, weights, os.path.join(REFERENCE_DIR, "final_model", "roomnet.meta"))
    r = o.forward(x)
    assert np.array_equal(out["dense_3/BiasAdd"], r["pre_relu6"])
    assert np.array_equal(out["Softmax"], r["softmax"]) and np.array_equal(out["ArgMax"], r["argmax"])
    from collections import Counter
    ops_run = Counter(op for op, _ in interp.executed_ops)
    assert (ops_run["Conv2D"], ops_run["AvgPool"], ops_run["FusedBatchNorm"], ops_run["ResizeBilinear"],
            ops_run["MatMul"], ops_run["Relu6"]) == (10, 9, 13, 3, 4, 14)


@needs_reference
def test_fixture_checkpoint_is_byte_identical_to_reference(ckpt_prefix):
    for ext in (".index", ".data-00000-of-00001"):
        a = open(os.path.join(REFERENCE_DIR, "final_model", "roomnet" + ext), "rb").read()
        assert a == open(ckpt_prefix + ext, "rb").read()
