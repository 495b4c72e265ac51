"""Row n4 probe: which output channels could be pruned EXACTLY?  (CPU only; uses the oracle's fp64 fold.)

A channel can be removed without changing any logit only if its pre-activation is provably constant over all inputs:
|sum_k w_k x_k| <= eps = sum_k |w_k| * max|x_k|.  SURVEY App. E counts channels with max|w| < 1e-4 as "dead"; this
prints, per conv layer, how many of those also have a worst-case excursion eps below 1e-6 - i.e. are safe to fold into
the consumer's bias - and the largest eps among the rest."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from oracle import fold as F, tf_bundle  # noqa: E402
from roomnet_b200.workload import default_checkpoint_prefix  # noqa: E402

w = tf_bundle.load_checkpoint(default_checkpoint_prefix())
f = F.fold(w)
bound = np.full(3, 255.0)
for i, L in enumerate(f["convs"]):
    W = L["W"]
    eps = (np.abs(W) * bound[None, None, :, None]).sum(axis=(0, 1, 2))
    small_w = np.abs(W).max(axis=(0, 1, 2)) < 1e-4
    safe = eps < 1e-6
    rest = eps[small_w & ~safe]
    print("conv%d  cout %3d  max|w|<1e-4: %3d  provably constant (eps<1e-6): %3d  worst eps of the others: %s"
          % (i, W.shape[3], small_w.sum(), safe.sum(), ("%.2e" % rest.max()) if rest.size else "-"))
    out_b = np.full(W.shape[3], 6.0)
    if i in f["joins"]:
        A, B, C, _ = f["joins"][i]
        out_b = np.abs(A) * 6 + np.abs(B) * 6 + np.abs(C)
    bound = out_b
