"""GPU check of the fused residual-block-2 kernel against the layer-by-layer kernels (same math, different rounding
points: agreement to 16-bit noise) and against the golden logits.  Usage: python tools/check_block2.py [n_images] [side]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from roomnet_b200 import _capi  # noqa: E402
from roomnet_b200.workload import default_checkpoint_prefix, synthetic_dense0, synthetic_suite  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
side = int(sys.argv[2]) if len(sys.argv) > 2 else 224
imgs = synthetic_suite(max(n, 1), side)[:n] if side != 224 else synthetic_suite(64)[np.arange(n) % 64]
outs = {}
for name, lw in (("layerwise", True), ("fused", False)):
    h = _capi.Handle(im_side=side, precision="fp16", layerwise=lw, max_batch=max(n, 1))
    if side != 224:
        h.set_dense0(synthetic_dense0(side))
    h.load_tf_checkpoint(default_checkpoint_prefix())
    t0 = time.time()
    top1, probs, logits = h.infer_u8_bgr(imgs, want_logits=True)
    print(name, "infer ok in %.3fs, launches %d" % (time.time() - t0, h.kernel_launches), flush=True)
    outs[name] = (top1, logits, h.debug_activation(3), h.debug_activation(1))
    h.close()
a3, b3 = outs["layerwise"][2], outs["fused"][2]
d = np.abs(a3 - b3)
print("layer3 (block output) max|diff| %.3e  mismatching elements %d / %d" % (d.max(), (d > 0).sum(), d.size))
if d.max() > 0:
    idx = np.argwhere(d > 0)
    print("first mismatches (n, y, x, c):", idx[:8].tolist())
    print("rows with mismatches:", np.unique(idx[:, 1])[:40].tolist())
    print("cols with mismatches:", np.unique(idx[:, 2])[:40].tolist())
print("logits max|diff| %.3e, top1 equal %s" % (np.abs(outs["layerwise"][1] - outs["fused"][1]).max(),
                                               np.array_equal(outs["layerwise"][0], outs["fused"][0])))
if side == 224:
    g = np.load(os.path.join(ROOT, "tests", "golden", "suite64.npz"))
    idx = np.arange(n) % 64
    print("fused vs golden: max|dlogit| %.3e top1 equal %s" % (np.abs(outs["fused"][1] - g["logits"][idx]).max(),
                                                              np.array_equal(outs["fused"][0], g["argmax"][idx])))
