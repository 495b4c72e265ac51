"""ORACLE / TEST INFRASTRUCTURE — not part of the product path.

CPU restatement of the reference's inference forward pass:

* ``RoomNet.init_nn_graph``   reference network.py:225-244
* ``RoomNet.conv_block``      reference network.py:172-208
* ``RoomNet.dense_block``     reference network.py:210-223
* ``RoomNet.infer``           reference network.py:128-135
* ``RoomNet.infer_optimized`` reference network.py:148-156
* ``RoomNet.center_crop``     reference network.py:137-146
* softmax / argmax            reference network.py:44-45

PARITY UNPINNED: the reference ships no tests, golden vectors or fixtures for
this path and TensorFlow 1.13.1 cannot be installed here, so this oracle cannot
be checked against the reference's own outputs.  What pins it instead:
(1) the checkpoint bytes (CRC32C of every tensor, sha256 of the files);
(2) ``oracle/tf_graph_interp.py`` executes the reference's *shipped graph*
(final_model/roomnet.meta) node by node and must agree with this hand-written
restatement (tests/test_oracle.py, runs where /root/reference is present);
(3) the fp32 and fp64 twins must agree to ~1e-4 on logits.

The graph is kept UNFOLDED (conv → ReLU6 → AvgPool → BN, residual through a
legacy bilinear resize, BN after the add) exactly as the reference builds it.
"""
from __future__ import annotations

import numpy as np

from . import tf_ops as ops
from .tf_bundle import default_checkpoint_prefix, load_checkpoint

CLASS_LABELS = ['Backyard', 'Bathroom', 'Bedroom', 'Frontyard', 'Kitchen', 'LivingRoom']  # reference infer.py:22

# (output_filters, pooling, pool_ksize, pool_stride, block_depth) — reference network.py:226-230
CONV_BLOCKS = [
    (8, True, 3, 1, 1),
    (32, True, 4, 1, 3),
    (64, True, 4, 2, 2),
    (128, False, 0, 0, 1),
    (16, True, 4, 2, 3),
]
DENSE_UNITS = [32, 16, 8]  # then num_classes, biased, no BN — reference network.py:234-237


def spatial_trace(im_side: int):
    """Spatial size after every conv/pool; returns list of dicts per conv."""
    s = im_side
    out = []
    for (_, pooling, k, st, depth) in CONV_BLOCKS:
        for _ in range(depth):
            conv = s - 2
            pooled = (conv - k) // st + 1 if pooling else conv
            out.append(dict(inp=s, conv=conv, out=pooled, k=k if pooling else 0, s=st if pooling else 0))
            s = pooled
    return out


def flat_len(im_side: int) -> int:
    return spatial_trace(im_side)[-1]["out"] ** 2 * CONV_BLOCKS[-1][0]


def _bn_name(i):
    return "batch_normalization" if i == 0 else "batch_normalization_%d" % i


def _conv_name(i):
    return "conv2d" if i == 0 else "conv2d_%d" % i


def _dense_name(i):
    return "dense" if i == 0 else "dense_%d" % i


class RoomNetOracle:
    """Mirror of the reference ``RoomNet`` (optimized_inference=True) on the CPU."""

    def __init__(self, num_classes=6, im_side=224, dtype=np.float32, weights=None,
                 dense0_kernel=None, conv_backend="numpy"):
        self.num_classes = num_classes
        self.im_side = im_side
        self.dtype = np.dtype(dtype)
        self.conv_backend = conv_backend
        self.weights = weights
        self.dense0_kernel = dense0_kernel

    # -- reference network.py:105-126 (only the explicit-path restore is on the hot path)
    def load(self, model_path=None):
        self.weights = load_checkpoint(model_path or default_checkpoint_prefix())
        return self

    def _w(self, name):
        return self.weights[name].astype(self.dtype)

    def _bn(self, x, idx):
        n = _bn_name(idx)
        return ops.batch_norm_inference(x, self._w(n + "/gamma"), self._w(n + "/beta"),
                                        self._w(n + "/moving_mean"), self._w(n + "/moving_variance"))

    @staticmethod
    def center_crop(x):  # reference network.py:137-146 (incl. the floor-division quirk)
        h, w, _ = x.shape
        offset = abs((w - h) // 2)
        if h < w:
            return x[:, offset:offset + h, :]
        elif w < h:
            return x[offset:offset + w, :, :]
        return x.copy()

    @staticmethod
    def normalise(im_bgr_u8):  # reference network.py:129 / :153 — float64 math, cast at feed
        return ((im_bgr_u8[..., [2, 1, 0]] / 255.) * 2) - 1

    def forward(self, x, collect=False):
        """``sess.run`` restated: x is NHWC float (RGB, [-1,1]); returns dict.

        keys: 'pre_relu6' (dense_3/BiasAdd), 'logits' (= out_op, ReLU6-clipped,
        reference network.py:43), 'softmax', 'argmax'; with collect=True also every
        block tensor keyed by the TF node name.
        """
        x = np.asarray(x).astype(self.dtype)  # feed cast float64→float32 (RNE)
        t = {} if collect else None
        bn_i = 0
        conv_i = 0
        out = x
        for (filters, pooling, k, st, depth) in CONV_BLOCKS:
            residual = None
            for d in range(depth):
                out = ops.conv2d_valid(out, self._w(_conv_name(conv_i) + "/kernel"), self.conv_backend)
                out = ops.relu6(out)
                if collect:
                    t[_conv_name(conv_i) + "/Relu6"] = out
                conv_i += 1
                if pooling:
                    out = ops.avg_pool_valid(out, k, st)
                out = self._bn(out, bn_i)
                if collect:
                    t[_bn_name(bn_i)] = out
                bn_i += 1
                if d == 0:
                    residual = out
            if depth > 1:
                out = out + ops.resize_bilinear_legacy(residual, out.shape[1], out.shape[2])
                out = self._bn(out, bn_i)
                if collect:
                    t[_bn_name(bn_i)] = out
                bn_i += 1
        n = out.shape[0]
        out = out.reshape(n, -1)  # NHWC flatten (h, w, c) — reference network.py:231-234
        for di, units in enumerate(DENSE_UNITS):
            if di == 0 and self.dense0_kernel is not None:
                kern = np.asarray(self.dense0_kernel).astype(self.dtype)
            else:
                kern = self._w(_dense_name(di) + "/kernel")
            out = out @ kern
            out = ops.relu6(out)
            out = self._bn(out, bn_i)
            bn_i += 1
        pre = out @ self._w("dense_3/kernel") + self._w("dense_3/bias")
        logits = ops.relu6(pre)
        sm = ops.softmax(logits)
        res = dict(pre_relu6=pre, logits=logits, softmax=sm, argmax=ops.argmax_first(sm))
        if collect:
            res["tensors"] = t
        return res

    # -- reference network.py:128-135
    def infer(self, im_in):
        im = self.normalise(np.asarray(im_in))
        return self.forward(im)["argmax"]

    # -- reference network.py:148-156
    def preprocess(self, im_in):
        import cv2
        im = self.center_crop(im_in)
        h, w, _ = im.shape
        if h != self.im_side or w != self.im_side:
            im = cv2.resize(im, (self.im_side, self.im_side))
        return im

    def infer_optimized(self, im_in):
        im = self.normalise(self.preprocess(im_in))
        im = np.expand_dims(im, 0)
        r = self.forward(im)
        return r["argmax"], r["softmax"]


def synthetic_image(seed: int, side: int = 224) -> np.ndarray:
    """Fixed synthetic BGR uint8 suite: family = seed % 4."""
    import cv2
    rng = np.random.default_rng(seed)
    fam = seed % 4
    if fam == 0:
        return rng.integers(0, 256, (side, side, 3), dtype=np.uint8)
    if fam == 1:
        k = [2, 4, 8, 16, 32, 64][(seed // 4) % 6]
        small = rng.integers(0, 256, (k, k, 3), dtype=np.uint8)
        return cv2.resize(small, (side, side), interpolation=cv2.INTER_CUBIC)
    if fam == 2:
        col = rng.integers(0, 256, 3, dtype=np.uint8)
        return np.broadcast_to(col, (side, side, 3)).copy()
    c0 = rng.integers(0, 256, 3).astype(np.float64)
    c1 = rng.integers(0, 256, 3).astype(np.float64)
    ramp = np.linspace(0.0, 1.0, side)
    g = c0[None, :] + (c1 - c0)[None, :] * ramp[:, None]  # [side, 3]
    if (seed // 4) % 2 == 0:
        img = np.broadcast_to(g[None, :, :], (side, side, 3))
    else:
        img = np.broadcast_to(g[:, None, :], (side, side, 3))
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


def synthetic_suite(n: int = 64, side: int = 224) -> np.ndarray:
    return np.stack([synthetic_image(s, side) for s in range(n)])


def synthetic_dense0(im_side: int) -> np.ndarray:
    """Synthetic dense/kernel for im_side != 224."""
    fl = flat_len(im_side)
    return np.random.default_rng(1234).normal(0, 0.02, (fl, 32)).astype(np.float32)
