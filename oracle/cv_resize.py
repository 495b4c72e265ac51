"""ORACLE / TEST INFRASTRUCTURE — not part of the product path.

Integer restatement of OpenCV's ``cv2.resize(img, (S, S))`` default path for uint8 images (INTER_LINEAR,
fixed-point, half-pixel centres, no antialiasing) — the call the reference makes at network.py:152 — and of
``center_crop`` (network.py:137-146).  OpenCV is a third-party dependency of the reference (unpinned); the algorithm
below follows its published resize.cpp: 11-bit horizontal/vertical coefficients (saturate_cast<short>(w*2048)),
int32 horizontal pass, and the vertical pass ((b0*(S0>>4))>>16) + ((b1*(S1>>4))>>16) + 2) >> 2; exact 2x
down-scaling in both directions is routed to the area filter ((a+b+c+d+2)>>2) as OpenCV does.
tests/test_preprocess.py pins it bit-exactly against the installed cv2 on random shapes.
"""
from __future__ import annotations

import numpy as np


def _coeffs(dst: int, src: int, vertical: bool = False):
    """Tap indices and 11-bit weights of one axis.

    Horizontal axis: out-of-range taps are folded by clamping the fraction (fx = 0 at the borders).
    Vertical axis: OpenCV keeps the fraction and only clips the two ROW INDICES, so a border row is blended
    with itself through two separately truncated products (this differs by one LSB from the clamped form).
    """
    scale = np.float64(src) / np.float64(dst)
    d = np.arange(dst, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int64)
    f = (f - s.astype(np.float32)).astype(np.float32)
    if vertical:
        s0 = np.clip(s, 0, src - 1)
        s1 = np.clip(s + 1, 0, src - 1)
    else:
        neg = s < 0
        f[neg] = 0
        s[neg] = 0
        hi = s >= src - 1
        f[hi] = 0
        s[hi] = src - 1
        s0 = s
        s1 = np.minimum(s + 1, src - 1)
    w1 = np.rint(f * np.float32(2048.0)).astype(np.int64)
    w0 = np.rint((np.float32(1.0) - f) * np.float32(2048.0)).astype(np.int64)
    return s0, s1, w0, w1


def resize_linear_u8(img: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    h, w = img.shape[:2]
    if (h, w) == (out_h, out_w):
        return img.copy()
    if h == 2 * out_h and w == 2 * out_w:  # OpenCV: INTER_LINEAR with integer scale 2x2 -> INTER_AREA fast path
        a = img.astype(np.int64)
        return ((a[0::2, 0::2] + a[0::2, 1::2] + a[1::2, 0::2] + a[1::2, 1::2] + 2) >> 2).astype(np.uint8)
    x0, x1, a0, a1 = _coeffs(out_w, w)
    y0, y1, b0, b1 = _coeffs(out_h, h, vertical=True)
    src = img.astype(np.int64)
    shp = (1, -1) + (1,) * (img.ndim - 2)
    hrow = src[:, x0] * a0.reshape(shp) + src[:, x1] * a1.reshape(shp)  # [h, out_w, c] int32 range
    s0, s1 = hrow[y0], hrow[y1]
    shp_b = (-1, 1) + (1,) * (img.ndim - 2)
    v = ((b0.reshape(shp_b) * (s0 >> 4)) >> 16) + ((b1.reshape(shp_b) * (s1 >> 4)) >> 16)
    return np.clip((v + 2) >> 2, 0, 255).astype(np.uint8)


def center_crop_rect(h: int, w: int):
    off = abs((w - h) // 2)  # Python floor division: the reference's rounding quirk for h > w
    if h < w:
        return 0, off, h
    if w < h:
        return off, 0, w
    return 0, 0, h


def preprocess(img: np.ndarray, side: int) -> np.ndarray:
    y0, x0, s = center_crop_rect(*img.shape[:2])
    return resize_linear_u8(img[y0:y0 + s, x0:x0 + s], side, side)
