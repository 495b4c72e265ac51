"""Runs a few forward passes of one micro-batch through the device-resident entry point (for ncu)."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from roomnet_b200.workload import default_checkpoint_prefix, synthetic_suite  # noqa: E402
from roomnet_b200 import _capi  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--iters", type=int, default=2)
ap.add_argument("--precision", default="fp16")
ap.add_argument("--side", type=int, default=224)
args = ap.parse_args()

h = _capi.Handle(precision=args.precision, im_side=args.side, max_batch=args.batch)
if args.side != 224:
    from roomnet_b200.workload import synthetic_dense0
    h.set_dense0(synthetic_dense0(args.side))
h.load_tf_checkpoint(default_checkpoint_prefix())
imgs = synthetic_suite(64, args.side)[np.arange(args.batch) % 64]
d_in = torch.from_numpy(np.ascontiguousarray(imgs)).cuda()
d_top1 = torch.empty(args.batch, dtype=torch.int64, device="cuda")
d_probs = torch.empty(args.batch, 6, device="cuda")
for _ in range(args.iters):
    h.infer_u8_bgr_device(d_in.data_ptr(), args.batch, d_top1.data_ptr(), d_probs.data_ptr(), None, None)
torch.cuda.synchronize()
print("top1", d_top1[:8].tolist(), "launches/iter", h.kernel_launches)
