"""TEST INFRASTRUCTURE (oracle): CPU restatement of the mobile module's camera front end, for the parity tests of
rn_infer_yuv420 only - never imported by the product.

Follows the reference Java sources:
  * ``yuv420_to_rgb``  - ImageUtils.convertYUV420ToARGB8888 (ImageUtils.java:131-151) with YUV2RGB (:100-129): integer
    arithmetic, restated exactly (kMaxChannelValue = 262143, :28).
  * ``frame_to_crop``  - ImageUtils.getTransformationMatrix (:168-225) as used by ClassifierActivity.java:89-106
    (MAINTAIN_ASPECT = true, :40) followed by ``canvas.drawBitmap(rgbFrameBitmap, frameToCropTransform, null)``: every
    crop pixel centre is mapped back through the inverse transform and takes the nearest frame pixel (no Paint -> no
    bitmap filtering).  Android itself cannot run here, so the sampling rule is a restatement of the documented
    behaviour, not a pinned one; the colour arithmetic is pinned by the Java source.
"""
import numpy as np

K_MAX = 262143  # ImageUtils.java:28


def yuv2rgb(y, u, v):
    """ImageUtils.YUV2RGB (:100-129) on int arrays; returns (r8, g8, b8) as the ARGB word packs them."""
    y = np.maximum(y.astype(np.int64) - 16, 0)
    u = u.astype(np.int64) - 128
    v = v.astype(np.int64) - 128
    y1192 = 1192 * y
    r = np.clip(y1192 + 1634 * v, 0, K_MAX)
    g = np.clip(y1192 - 833 * v - 400 * u, 0, K_MAX)
    b = np.clip(y1192 + 2066 * u, 0, K_MAX)
    # 0xff000000 | ((r << 6) & 0xff0000) | ((g >> 2) & 0xff00) | ((b >> 10) & 0xff)
    return ((r << 6) & 0xFF0000) >> 16, ((g >> 2) & 0xFF00) >> 8, (b >> 10) & 0xFF


def yuv420_to_rgb(y, u, v, width, height, y_row_stride, uv_row_stride, uv_pixel_stride):
    """convertYUV420ToARGB8888 (:131-151): full frame, [height, width, 3] uint8 in R,G,B order."""
    jj, ii = np.meshgrid(np.arange(height), np.arange(width), indexing="ij")
    uv = uv_row_stride * (jj >> 1) + (ii >> 1) * uv_pixel_stride
    r, g, b = yuv2rgb(y[y_row_stride * jj + ii], u[uv], v[uv])
    return np.stack([r, g, b], axis=-1).astype(np.uint8)


def frame_to_crop(rgb, side, rotation):
    """getTransformationMatrix(w, h, side, side, rotation, True) + unfiltered drawBitmap: [side, side, 3]."""
    height, width = rgb.shape[:2]
    rotation %= 360
    transpose = (rotation + 90) % 180 == 0
    in_w, in_h = (height, width) if transpose else (width, height)
    s = 1.0
    if in_w != side or in_h != side:
        s = max(side / in_w, side / in_h)
    dy, dx = np.meshgrid(np.arange(side), np.arange(side), indexing="ij")
    px, py = dx + 0.5, dy + 0.5
    if rotation != 0:
        px, py = px - side / 2.0, py - side / 2.0
    px, py = px / s, py / s
    if rotation == 90:
        qx, qy = py, -px
    elif rotation == 180:
        qx, qy = -px, -py
    elif rotation == 270:
        qx, qy = -py, px
    else:
        qx, qy = px, py
    if rotation != 0:
        qx, qy = qx + width / 2.0, qy + height / 2.0
    i = np.clip(np.floor(qx).astype(np.int64), 0, width - 1)
    j = np.clip(np.floor(qy).astype(np.int64), 0, height - 1)
    return rgb[j, i]


def camera_front(y, u, v, width, height, y_row_stride, uv_row_stride, uv_pixel_stride, side, rotation):
    return frame_to_crop(yuv420_to_rgb(y, u, v, width, height, y_row_stride, uv_row_stride, uv_pixel_stride), side,
                         rotation)
