"""Synthetic workload of the benchmark and the profiling tools (product side: no dependency on oracle/).

Topology constants of the reference graph (network.py:226-237), the spatial trace they imply, the fixed synthetic
image suite of SURVEY §8d and the location of the shipped ``final_model`` weights.  ``tests/test_host.py`` pins the
generators against the oracle's own copies (same seeds -> identical bytes).
"""
from __future__ import annotations

import os

import numpy as np

# (output_filters, pooling, pool_ksize, pool_stride, block_depth) — reference network.py:226-230
CONV_BLOCKS = [(8, True, 3, 1, 1), (32, True, 4, 1, 3), (64, True, 4, 2, 2), (128, False, 0, 0, 1), (16, True, 4, 2, 3)]
JOIN_SOURCE = {3: 1, 5: 4, 9: 7}  # conv index of a block's last layer -> conv index whose pooled output is the residual


def channels():
    """Channel count before conv i (index 0 = the image) for the 10 convolutions."""
    ch = [3]
    for (filters, _, _, _, depth) in CONV_BLOCKS:
        ch += [filters] * depth
    return ch


def spatial_trace(im_side: int):
    """Per conv: input side, conv output side (3x3 VALID), pooled side, pool window and stride."""
    s, out = im_side, []
    for (_, pooling, k, st, depth) in CONV_BLOCKS:
        for _ in range(depth):
            conv = s - 2
            pooled = (conv - k) // st + 1 if pooling else conv
            out.append(dict(inp=s, conv=conv, out=pooled, k=k if pooling else 0, s=st if pooling else 0))
            s = pooled
    return out


def flat_len(im_side: int) -> int:
    return spatial_trace(im_side)[-1]["out"] ** 2 * CONV_BLOCKS[-1][0]


def conv_flops(layer: int, side: int = 224) -> int:
    """Algorithmic FLOPs (2 x MAC) of conv `layer` for one image."""
    t, ch = spatial_trace(side)[layer], channels()
    return 2 * t["conv"] * t["conv"] * ch[layer + 1] * 9 * ch[layer]


def conv_bytes(layer: int, side: int = 224) -> int:
    """Algorithmic HBM bytes of one image through one layer kernel of the 16-bit path: input tensor + output tensor
    (+ the residual source of the fused joins), 2 bytes per element; conv0's input is the 16-byte-per-pixel pair-chunk
    tensor written by prep_u8."""
    tr, ch = spatial_trace(side), channels()
    t = tr[layer]
    b = t["inp"] ** 2 * (16 if layer == 0 else 2 * ch[layer]) + t["out"] ** 2 * 2 * ch[layer + 1]
    src = JOIN_SOURCE.get(layer)
    if src is not None:
        b += tr[src]["out"] ** 2 * 2 * ch[src + 1]
    return b


def synthetic_image(seed: int, side: int = 224) -> np.ndarray:
    """BGR uint8 image of the fixed suite: family = seed % 4 — uniform noise, low-frequency noise
    (cubic upsampling of a k x k grid), solid colour, linear gradient between two colours."""
    import cv2
    rng = np.random.default_rng(seed)
    family = seed % 4
    if family == 0:
        return rng.integers(0, 256, (side, side, 3), dtype=np.uint8)
    if family == 1:
        k = [2, 4, 8, 16, 32, 64][(seed // 4) % 6]
        return cv2.resize(rng.integers(0, 256, (k, k, 3), dtype=np.uint8), (side, side), interpolation=cv2.INTER_CUBIC)
    if family == 2:
        return np.broadcast_to(rng.integers(0, 256, 3, dtype=np.uint8), (side, side, 3)).copy()
    first = rng.integers(0, 256, 3).astype(np.float64)
    second = rng.integers(0, 256, 3).astype(np.float64)
    line = first[None, :] + (second - first)[None, :] * np.linspace(0.0, 1.0, side)[:, None]
    img = line[None, :, :] if (seed // 4) % 2 == 0 else line[:, None, :]
    return np.clip(np.rint(np.broadcast_to(img, (side, side, 3))), 0, 255).astype(np.uint8)


def synthetic_suite(n: int = 64, side: int = 224) -> np.ndarray:
    return np.stack([synthetic_image(seed, side) for seed in range(n)])


def synthetic_dense0(im_side: int) -> np.ndarray:
    """dense/kernel for im_side != 224 (BASELINE config 4): the shipped [64, 32] kernel does not fit."""
    return np.random.default_rng(1234).normal(0, 0.02, (flat_len(im_side), 32)).astype(np.float32)


def default_checkpoint_prefix() -> str:
    """The shipped ``final_model`` weights: the byte-identical copy under final_model/ at the repository root (the
    reference's own layout, infer.py:24; the GPU box has no /root/reference), else the reference tree of the build
    container."""
    here = os.path.dirname(os.path.abspath(__file__))
    fixture = os.path.join(here, "..", "final_model", "roomnet")
    if os.path.exists(fixture + ".index"):
        return os.path.normpath(fixture)
    return "/root/reference/final_model/roomnet"
