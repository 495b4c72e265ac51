"""Per-layer activation error of the 16-bit path against the folded fp64 oracle (test infrastructure import, tools only).
Usage: python tools/layer_errors.py [n_images] [fp16|bf16|fp32]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import tf_bundle  # noqa: E402
from oracle.fold import fold, folded_forward  # noqa: E402
from roomnet_b200 import _capi  # noqa: E402
from roomnet_b200.workload import default_checkpoint_prefix, synthetic_suite  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
precision = sys.argv[2] if len(sys.argv) > 2 else "fp16"
imgs = synthetic_suite(64)[:n]
weights = tf_bundle.load_checkpoint(default_checkpoint_prefix())
ref = folded_forward(fold(weights), imgs, dtype=np.float64, conv_backend="torch", collect=True)
for lw in (True, False):
    h = _capi.Handle(precision=precision, layerwise=lw, max_batch=max(n, 1))
    h.load_tf_checkpoint(default_checkpoint_prefix())
    t, p, l = h.infer_u8_bgr(imgs, want_logits=True)
    errs = []
    for layer in range(10):
        try:
            got = h.debug_activation(layer)
        except _capi.RoomNetError:
            errs.append("  -  ")
            continue
        want = ref["tensors"][layer]
        errs.append("%.1e" % (np.abs(got - want).max() / (np.abs(want).max() + 1e-6)))
    print("layerwise" if lw else "fused    ", " ".join(errs), "| max|dlogit| %.2e" % np.abs(l - ref["logits"]).max())
    h.close()
