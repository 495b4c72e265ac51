"""CPU-side cost of enqueuing one forward pass (no synchronisation inside the timed loop)."""
import os, sys, time
import numpy as np
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from roomnet_b200.workload import default_checkpoint_prefix, synthetic_suite  # noqa: E402
from roomnet_b200 import _capi  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
h = _capi.Handle(precision="fp16", max_batch=256)
h.load_tf_checkpoint(default_checkpoint_prefix())
imgs = synthetic_suite(64)[np.arange(n) % 64]
d_in = torch.from_numpy(np.ascontiguousarray(imgs)).cuda()
d_top1 = torch.empty(n, dtype=torch.int64, device="cuda")
d_probs = torch.empty(n, 6, device="cuda")
for _ in range(5):
    h.infer_u8_bgr_device(d_in.data_ptr(), n, d_top1.data_ptr(), d_probs.data_ptr(), None, None)
torch.cuda.synchronize()
for reps in (1, 20):
    t0 = time.perf_counter()
    for _ in range(reps):
        h.infer_u8_bgr_device(d_in.data_ptr(), n, d_top1.data_ptr(), d_probs.data_ptr(), None, None)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"n={n} reps={reps}: enqueue {1e6 * (t1 - t0) / reps:.1f} us/call, drain {1e6 * (t2 - t1):.1f} us")
