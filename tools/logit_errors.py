"""Per-image logit error of the 16-bit path against the golden fp64 logits: which images carry the error?"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from roomnet_b200 import _capi  # noqa: E402
from roomnet_b200.workload import default_checkpoint_prefix, synthetic_suite  # noqa: E402

g = np.load(os.path.join(ROOT, "tests", "golden", "suite64.npz"))
imgs = synthetic_suite(64)
for lw in (False, True):
    h = _capi.Handle(precision="fp16", layerwise=lw, max_batch=64)
    h.load_tf_checkpoint(default_checkpoint_prefix())
    t, p, l = h.infer_u8_bgr(imgs, want_logits=True)
    e = np.abs(l - g["logits"]).max(axis=1)
    order = np.argsort(-e)[:8]
    fam = ["noise", "lowfreq", "flat", "gradient"]
    print("layerwise" if lw else "fused    ", "max %.2e mean %.2e | worst:" % (e.max(), e.mean()),
          ", ".join("#%d(%s) %.1e" % (i, fam[i % 4], e[i]) for i in order))
    print("   per family max:", {fam[k]: "%.1e" % e[k::4].max() for k in range(4)}, "top1 ok", bool((t == g["argmax"]).all()))
    h.close()
