"""Runs the front-end kernels once each (for ncu): per-image crop+resize, batched crop+resize, JPEG decode (inverse DCT,
upsampling + colour), YUV camera front end,
ARGB feed, 300x300 and 600x600 forward passes (their tails use the unfused fp32 kernels)."""
import os
import sys

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
from roomnet_b200 import _capi  # noqa: E402
from roomnet_b200.workload import default_checkpoint_prefix, synthetic_dense0, synthetic_suite  # noqa: E402

rng = np.random.default_rng(0)
h = _capi.Handle(precision="fp16", max_batch=16)
h.load_tf_checkpoint(default_checkpoint_prefix())
photos = [rng.integers(0, 256, (int(rng.integers(300, 900)), int(rng.integers(300, 900)), 3), dtype=np.uint8)
          for _ in range(12)]
print("preprocess", h.preprocess_u8(photos[0]).shape)
print("images", h.infer_images_u8_bgr(photos)[0].tolist())
W, H = 640, 480
y = rng.integers(0, 256, W * H, dtype=np.uint8)
uv = rng.integers(0, 256, W * (H // 2) + 1, dtype=np.uint8)
print("yuv", h.infer_yuv420(y, uv[1:], uv[:-1], W, H, W, W, 2, 90)[0].tolist())
import cv2  # noqa: E402
jpegs = [cv2.imencode(".jpg", cv2.resize(p, None, fx=3, fy=3, interpolation=cv2.INTER_CUBIC),
                      [cv2.IMWRITE_JPEG_QUALITY, 90])[1].tobytes() for p in photos[:8]]
print("jpeg", h.infer_jpeg(jpegs)[0].tolist())
print("argb", h.infer_argb8888(rng.integers(0, 2 ** 31, (2, 224, 224), dtype=np.int64).astype(np.int32))[0].tolist())
h.close()
for side in (300, 600):
    hs = _capi.Handle(im_side=side, precision="fp16", max_batch=2)
    hs.set_dense0(synthetic_dense0(side))
    hs.load_tf_checkpoint(default_checkpoint_prefix())
    print(side, hs.infer_u8_bgr(synthetic_suite(2, side))[0].tolist())
    hs.close()
