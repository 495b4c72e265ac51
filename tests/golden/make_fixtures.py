"""Generates tests/golden/suite64.npz — run in the build container (needs /root/reference).

The vectors come from executing the reference's SHIPPED graph (final_model/roomnet.meta,
TensorFlow-1.13.1 MetaGraphDef) with oracle/tf_graph_interp.py in float64 on the fixed
synthetic suite.  They are NOT TensorFlow outputs — TensorFlow cannot be
installed here — so parity stays "unpinned" (see oracle/roomnet_oracle.py header); they pin
the oracle restatement, the product and future refactors to the shipped graph + checkpoint.
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)

from oracle.roomnet_oracle import RoomNetOracle, synthetic_suite  # noqa: E402
from oracle.tf_bundle import load_checkpoint  # noqa: E402
from oracle.tf_graph_interp import run_reference_graph  # noqa: E402

REF = "/root/reference/final_model/roomnet"


def main():
    imgs = synthetic_suite(64)
    weights = load_checkpoint(REF)
    x = RoomNetOracle.normalise(imgs)
    pre, logits, sm, am = [], [], [], []
    for i in range(0, 64, 8):
        out, _ = run_reference_graph(x[i:i + 8], weights, REF + ".meta", dtype=np.float64)
        pre.append(out["dense_3/BiasAdd"]); logits.append(out["Relu6_3"])
        sm.append(out["Softmax"]); am.append(out["ArgMax"])
        print("batch", i, out["ArgMax"])
    np.savez_compressed(
        os.path.join(os.path.dirname(__file__), "suite64.npz"),
        pre_relu6=np.concatenate(pre), logits=np.concatenate(logits), softmax=np.concatenate(sm),
        argmax=np.concatenate(am),
        image_md5=np.array([hashlib.md5(im.tobytes()).hexdigest() for im in imgs]),
    )


if __name__ == "__main__":
    main()
