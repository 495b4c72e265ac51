"""Where does the end-to-end rate go?  Times the rn_submit_u8_bgr / rn_wait pipeline at several depths and batch sizes
next to the device-resident call and the raw pinned-memory copy rate of the same box (run under gpurun)."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from roomnet_b200 import _capi  # noqa: E402
from roomnet_b200.workload import default_checkpoint_prefix  # noqa: E402

B = 256
STEPS = 200
dev = torch.device("cuda:0")
h = _capi.Handle(precision="fp16", max_batch=B)
h.load_tf_checkpoint(default_checkpoint_prefix())
rng = np.random.default_rng(0)
host = [torch.from_numpy(rng.integers(0, 256, (1024, 224, 224, 3), dtype=np.uint8)).pin_memory() for _ in range(4)]
devin = [x[:B].to(dev) for x in host]
d_top1 = torch.empty(B, dtype=torch.int64, device=dev)
d_probs = torch.empty(B, 6, device=dev)


def copy_rate():
    s = torch.cuda.Stream()
    d = torch.empty_like(devin[0])
    with torch.cuda.stream(s):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(3):
            d.copy_(host[0][:B], non_blocking=True)
        e0.record(s)
        for i in range(20):
            d.copy_(host[i % 4][:B], non_blocking=True)
        e1.record(s)
        s.synchronize()
    return d.numel() * 20 / e0.elapsed_time(e1) / 1e6


def device_rate():
    st = torch.cuda.current_stream().cuda_stream
    for i in range(5):
        h.infer_u8_bgr_device(devin[i % 4].data_ptr(), B, d_top1.data_ptr(), d_probs.data_ptr(), None, st)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(STEPS):
        h.infer_u8_bgr_device(devin[i % 4].data_ptr(), B, d_top1.data_ptr(), d_probs.data_ptr(), None, st)
    torch.cuda.synchronize()
    return B * STEPS / (time.perf_counter() - t0)


def pipelined(depth, nb=B, steps=STEPS):
    outs = [(torch.empty(nb, dtype=torch.int64).pin_memory(), torch.empty(nb, 6).pin_memory()) for _ in range(depth)]

    def run(k):
        tickets = []
        for i in range(k):
            t1, pr = outs[i % depth]
            tickets.append(h.submit_raw(host[i % 4].data_ptr(), nb, t1.data_ptr(), pr.data_ptr(), None))
            if len(tickets) >= depth:
                h.wait(tickets[-depth])
        h.wait(0)

    run(6)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    run(steps)
    torch.cuda.synchronize()
    return nb * steps / (time.perf_counter() - t0)


def submit_cost(nb=B):
    """host time of one rn_submit_u8_bgr with nothing to wait for"""
    t1 = torch.empty(nb, dtype=torch.int64).pin_memory()
    ts = []
    for i in range(20):
        h.wait(0)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        h.submit_raw(host[i % 4].data_ptr(), nb, t1.data_ptr(), None, None)
        ts.append(time.perf_counter() - t0)
    h.wait(0)
    return 1e6 * float(np.median(ts))


if __name__ == "__main__":
    print("pinned H2D copy rate: %.1f GB/s  (= %.0f k images/s of 224x224x3 bytes)" % (copy_rate(), copy_rate() * 1e6 / 150528))
    print("device-resident: %.1f k img/s" % (device_rate() / 1e3))
    print("host time of one submit (256 images): %.0f us" % submit_cost())
    for depth in (1, 2, 3, 4, 6):
        print("pipelined depth %d: %.1f k img/s" % (depth, pipelined(depth) / 1e3))
    for nb in (128, 512, 1024):
        print("pipelined depth 3, %d images per call: %.1f k img/s" % (nb, pipelined(3, nb, max(20, STEPS * B // nb)) / 1e3))
