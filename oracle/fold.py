"""ORACLE / TEST INFRASTRUCTURE — not part of the product path.

fp64 restatement of the exact forward BatchNorm fold and a
folded-form forward pass.  Used to (a) check the product's C++ folder
(roomnet_b200/csrc/fold.cpp) tensor by tensor, (b) prove the fold is exact
against the unfolded oracle, and (c) emulate reduced-precision operand rounding
(bf16 / split-bf16) on the CPU when choosing per-layer kernel precision.

Folded layer form:  p' = pool(relu6(conv_{W~}(p) + b~));  residual join
out = A*p_k + B*resize(p_0) + C  (reference network.py:184-203 restructured).
"""
from __future__ import annotations

import numpy as np

from . import tf_ops as ops
from .roomnet_oracle import CONV_BLOCKS, DENSE_UNITS, _bn_name, _conv_name, _dense_name

EPS = float(np.float32(0.0010000000474974513))


def bn_affine(w, idx):
    """s = gamma*rsqrt(var+eps), t = beta - mean*s in fp64."""
    n = _bn_name(idx)
    g = w[n + "/gamma"].astype(np.float64)
    b = w[n + "/beta"].astype(np.float64)
    m = w[n + "/moving_mean"].astype(np.float64)
    v = w[n + "/moving_variance"].astype(np.float64)
    s = g / np.sqrt(v + EPS)
    return s, b - m * s


def fold(weights: dict, dense0_kernel=None, u8_bgr_input=True):
    """Returns a dict describing the folded network (all fp64).

    convs: list of 10 dicts {W [3,3,Cin,Cout], b [Cout], pool_k, pool_s}
    joins: {conv index of last conv in block: (A, B, C, source conv index)}
    dense: list of 4 dicts {W [in,out], b [out]}
    """
    convs = []
    joins = {}
    bn_i = 0
    conv_i = 0
    pending = None  # (s, t) of the BN feeding the next conv/dense; None = input is final
    for (filters, pooling, k, st, depth) in CONV_BLOCKS:
        first = None
        for d in range(depth):
            W = weights[_conv_name(conv_i) + "/kernel"].astype(np.float64)
            if conv_i == 0:
                if u8_bgr_input:
                    # x_rgb[c] = v_bgr[2-c]*(2/255) - 1  (reference network.py:129/153)
                    b = -W.sum(axis=(0, 1, 2))
                    W = W[:, :, ::-1, :] * (2.0 / 255.0)
                else:
                    b = np.zeros(W.shape[3])
            elif pending is not None:
                s, t = pending
                b = np.einsum("hwio,i->o", W, t)
                W = W * s[None, None, :, None]
            else:
                b = np.zeros(W.shape[3])
            convs.append(dict(W=W, b=b, pool_k=k if pooling else 0, pool_s=st if pooling else 0))
            pending = bn_affine(weights, bn_i)
            if d == 0:
                first = (conv_i, pending)
            last = pending
            bn_i += 1
            conv_i += 1
        if depth > 1:
            s_r, t_r = bn_affine(weights, bn_i)
            bn_i += 1
            s_k, t_k = last
            src_conv, (s_0, t_0) = first
            joins[conv_i - 1] = (s_r * s_k, s_r * s_0, s_r * (t_k + t_0) + t_r, src_conv)
            pending = None
    dense = []
    names = [_dense_name(i) for i in range(4)]
    for di, nm in enumerate(names):
        if di == 0 and dense0_kernel is not None:
            W = np.asarray(dense0_kernel).astype(np.float64)
        else:
            W = weights[nm + "/kernel"].astype(np.float64)
        b = np.zeros(W.shape[1])
        if pending is not None:
            s, t = pending
            b = t @ W
            W = W * s[:, None]
        if nm + "/bias" in weights:
            b = b + weights[nm + "/bias"].astype(np.float64)
        dense.append(dict(W=W, b=b))
        if di < 3:
            pending = bn_affine(weights, bn_i)
            bn_i += 1
    return dict(convs=convs, joins=joins, dense=dense)


def round_bf16(x):
    """Round-to-nearest-even to bfloat16, returned in the input float dtype."""
    a = np.ascontiguousarray(x, dtype=np.float32)
    u = a.view(np.uint32)
    r = ((u >> 16) & 1) + 0x7FFF
    out = ((u + r) & 0xFFFF0000).astype(np.uint32).view(np.float32)
    return out.astype(np.asarray(x).dtype) if np.asarray(x).dtype != np.float32 else out


def folded_forward(folded, x, dtype=np.float64, prec=None, conv_backend="numpy", collect=False):
    """Folded forward.  ``x``: NHWC uint8 BGR (or float RGB if folded with u8_bgr_input=False).

    prec: optional {conv index: mode}; mode in
      'bf16'   – both operands rounded to bf16, wide accumulate
      'a2'     – activation split hi+lo (2 products), bf16 weights
      'split3' – hi*hi + lo*hi + hi*lo
      'f32'/None – operands unrounded
    plus key 'store' → 'bf16' to round every stored pooled activation / join output.
    """
    dt = np.dtype(dtype).type
    prec = prec or {}
    store = prec.get("store")
    act = np.asarray(x).astype(dtype)
    pooled = {}
    tensors = {}
    for i, L in enumerate(folded["convs"]):
        W = L["W"].astype(dtype)
        b = L["b"].astype(dtype)
        mode = prec.get(i)

        def cv(a, w):
            return ops.conv2d_valid(np.ascontiguousarray(a), np.ascontiguousarray(w), conv_backend)

        if mode == "bf16":
            y = cv(round_bf16(act), round_bf16(W))
        elif mode == "a2":
            ah = round_bf16(act)
            al = round_bf16(act - ah)
            wh = round_bf16(W)
            y = cv(ah, wh) + cv(al, wh)
        elif mode == "split3":
            ah = round_bf16(act)
            al = round_bf16(act - ah)
            wh = round_bf16(W)
            wl = round_bf16(W - wh)
            y = cv(ah, wh) + cv(al, wh) + cv(ah, wl)
        else:
            y = cv(act, W)
        y = ops.relu6(y + b)
        if L["pool_k"]:
            y = ops.avg_pool_valid(y, L["pool_k"], L["pool_s"])
        if store == "bf16" and i >= prec.get("store_from", 0):
            y = round_bf16(y)
        pooled[i] = y
        if i in folded["joins"]:
            A, B, C, src = folded["joins"][i]
            r = ops.resize_bilinear_legacy(pooled[src], y.shape[1], y.shape[2])
            y = y * A.astype(dtype) + r * B.astype(dtype) + C.astype(dtype)
            if store == "bf16":
                y = round_bf16(y)
        if collect:
            tensors[i] = y
        act = y
    out = act.reshape(act.shape[0], -1)
    pre = None
    for di, D in enumerate(folded["dense"]):
        pre = out @ D["W"].astype(dtype) + D["b"].astype(dtype)
        out = ops.relu6(pre)
    sm = ops.softmax(out)
    res = dict(pre_relu6=pre, logits=out, softmax=sm, argmax=ops.argmax_first(sm))
    if collect:
        res["tensors"] = tensors
    return res
