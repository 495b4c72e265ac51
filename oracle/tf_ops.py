"""ORACLE / TEST INFRASTRUCTURE — not part of the product path.

NumPy restatements of the TensorFlow-1.13.1 ops the reference's inference
graph uses (op list taken from final_model/roomnet.meta; call sites in the
reference: network.py:184 conv2d+relu6, :189 avg_pool, :193/:201/:217
batch_normalization, :199 resize_bilinear, :212 dense, :214 relu6, :44 softmax,
:45 argmax).  TensorFlow is a third-party dependency that is absent from
/root/reference and not installable here, so each function restates the
published semantics of the op (NHWC, VALID, legacy ``align_corners=False``
bilinear without half-pixel centres, FusedBatchNorm inference with
epsilon=1e-3).  All functions are dtype-generic: feed float32 for the fp32
oracle, float64 for its high-precision twin.

An optional torch-CPU backend for ``conv2d_valid`` is used for *speed only*
(cpu_baseline timing and large parity batches); tests check that it agrees with
the NumPy implementation.
"""
from __future__ import annotations

import numpy as np

BN_EPS = np.float32(0.0010000000474974513)  # attr "epsilon" of every FusedBatchNorm node


def conv2d_valid(x: np.ndarray, w: np.ndarray, backend: str = "numpy") -> np.ndarray:
    """TF Conv2D, NHWC, strides 1, dilation 1, padding VALID, HWIO kernel.

    y[n,i,j,o] = sum_{di,dj,c} x[n,i+di,j+dj,c] * w[di,dj,c,o]  (cross-correlation).
    """
    kh, kw, cin, cout = w.shape
    n, h, wd, c = x.shape
    assert c == cin
    if backend == "torch":
        import torch
        import torch.nn.functional as F
        xt = torch.from_numpy(np.ascontiguousarray(x)).permute(0, 3, 1, 2)
        wt = torch.from_numpy(np.ascontiguousarray(w)).permute(3, 2, 0, 1)
        y = F.conv2d(xt, wt)
        return np.ascontiguousarray(y.permute(0, 2, 3, 1).numpy())
    oh, ow = h - kh + 1, wd - kw + 1
    y = np.zeros((n, oh, ow, cout), dtype=x.dtype)
    for di in range(kh):
        for dj in range(kw):
            patch = x[:, di:di + oh, dj:dj + ow, :]
            y += np.tensordot(patch, w[di, dj], axes=([3], [0]))
    return y


def relu6(x: np.ndarray) -> np.ndarray:
    return np.minimum(np.maximum(x, x.dtype.type(0)), x.dtype.type(6))


def avg_pool_valid(x: np.ndarray, k: int, s: int) -> np.ndarray:
    """TF AvgPool, NHWC, ksize k×k, stride s, VALID (all windows full → /k²)."""
    n, h, w, c = x.shape
    oh, ow = (h - k) // s + 1, (w - k) // s + 1
    acc = np.zeros((n, oh, ow, c), dtype=x.dtype)
    for di in range(k):
        for dj in range(k):
            acc += x[:, di:di + (oh - 1) * s + 1:s, dj:dj + (ow - 1) * s + 1:s, :]
    return acc / x.dtype.type(k * k)


def batch_norm_inference(x, gamma, beta, mean, var, eps=BN_EPS):
    """FusedBatchNorm(is_training=False): (x-mean)*rsqrt(var+eps)*gamma+beta."""
    dt = x.dtype.type
    inv = dt(1) / np.sqrt(var.astype(x.dtype) + dt(eps))
    scale = inv * gamma.astype(x.dtype)
    return x * scale + (beta.astype(x.dtype) - mean.astype(x.dtype) * scale)


def resize_bilinear_legacy(x: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """TF-1.13 ResizeBilinear, align_corners=False, no half-pixel centres.

    scale = in/out; src = dst*scale; lo = floor(src); hi = min(lo+1, in-1);
    top = tl + (tr-tl)*tx; bottom = bl + (br-bl)*tx; out = top + (bottom-top)*ty.
    Interpolation weights are computed in float32 as in the TF kernel.
    """
    n, h, w, c = x.shape
    dt = x.dtype.type

    def taps(in_size, out_size):
        scale = np.float32(in_size) / np.float32(out_size)
        src = np.arange(out_size, dtype=np.float32) * scale
        lo = np.floor(src).astype(np.int64)
        hi = np.minimum(lo + 1, in_size - 1)
        t = (src - lo.astype(np.float32)).astype(np.float32)
        return lo, hi, t

    ylo, yhi, ty = taps(h, out_h)
    xlo, xhi, tx = taps(w, out_w)
    tx = tx.astype(x.dtype)[None, None, :, None]
    ty = ty.astype(x.dtype)[None, :, None, None]
    top_l = x[:, ylo][:, :, xlo]
    top_r = x[:, ylo][:, :, xhi]
    bot_l = x[:, yhi][:, :, xlo]
    bot_r = x[:, yhi][:, :, xhi]
    top = top_l + (top_r - top_l) * tx
    bot = bot_l + (bot_r - bot_l) * tx
    out = top + (bot - top) * ty
    return out.astype(dt)


def softmax(x: np.ndarray) -> np.ndarray:
    m = x.max(axis=-1, keepdims=True)
    e = np.exp(x - m)
    return e / e.sum(axis=-1, keepdims=True)


def argmax_first(x: np.ndarray) -> np.ndarray:
    """tf.argmax: index of the first maximum, int64."""
    return np.argmax(x, axis=-1).astype(np.int64)
