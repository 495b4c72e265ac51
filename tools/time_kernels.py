"""Per-kernel device times (CUDA events between launches) for one batch, without the bench's correctness guard."""
import os, sys
import numpy as np
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import torch
from roomnet_b200.workload import default_checkpoint_prefix, synthetic_suite
from roomnet_b200 import _capi
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
precision = sys.argv[2] if len(sys.argv) > 2 else "fp16"
steps = 10
h = _capi.Handle(precision=precision); h.load_tf_checkpoint(default_checkpoint_prefix())
imgs = synthetic_suite(64)[np.arange(B) % 64]
d_in = torch.from_numpy(np.ascontiguousarray(imgs)).cuda()
d_top1 = torch.empty(B, dtype=torch.int64, device="cuda"); d_probs = torch.empty(B, 6, device="cuda")
run = lambda: h.infer_u8_bgr_device(d_in.data_ptr(), B, d_top1.data_ptr(), d_probs.data_ptr(), None, None)
for _ in range(3): run()
torch.cuda.synchronize()
h.set_profiling(True)
for _ in range(steps): run()
torch.cuda.synchronize()
prof = h.get_profile()
print({p["name"]: round(p["ms"] / steps, 4) for p in prof}, "total", round(sum(p["ms"] for p in prof) / steps, 3))
