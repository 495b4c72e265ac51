"""Measures the BASELINE.json parity/benchmark configs other than the headline one (bench.py covers configs[1]):

  config 3: batch 8192 sharded contiguously over g GPUs by ONE multi-device handle, logits gathered into one
            pinned host buffer; img/s and bit-identity of the logits across g
  config 4: im_side 300 and 600 (synthetic dense/kernel), batch 512, img/s (device-resident) + parity on 8 images
  config 5: batch-1 latency through the C ABI (the JNI shim adds < 1 us on top; see tests/jni_harness.cpp)
Prints one JSON object per config.
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from oracle.roomnet_oracle import RoomNetOracle, synthetic_dense0, synthetic_suite  # noqa: E402
from oracle.tf_bundle import default_checkpoint_prefix, load_checkpoint  # noqa: E402
from roomnet_b200 import _capi  # noqa: E402

ckpt = default_checkpoint_prefix()
ngpu = torch.cuda.device_count()
which = sys.argv[1:] or ["3", "4", "5"]

if "3" in which:
    suite = synthetic_suite(64)
    n = 8192
    x = torch.from_numpy(np.ascontiguousarray(suite[np.arange(n) % 64])).pin_memory()
    top1 = torch.empty(n, dtype=torch.int64).pin_memory()
    probs = torch.empty(n, 6).pin_memory()
    logits = torch.empty(n, 6).pin_memory()
    ref_logits = None
    for g in [1, 2, 4, 8]:
        if g > ngpu:
            break
        h = _capi.Handle(precision="fp16", devices=tuple(range(g)))
        h.load_tf_checkpoint(ckpt)
        run = lambda: h.infer_raw("rn_infer_u8_bgr", x.data_ptr(), n, top1.data_ptr(), probs.data_ptr(), logits.data_ptr())
        run()
        t0 = time.perf_counter()
        for _ in range(3):
            run()
        dt = (time.perf_counter() - t0) / 3
        l = logits.numpy().copy()
        if ref_logits is None:
            ref_logits = l
        print(json.dumps({"config": 3, "gpus": g, "batch": n, "images_per_s_e2e": n / dt,
                          "bit_identical_to_g1": bool(np.array_equal(l, ref_logits))}), flush=True)
        h.close()

if "4" in which:
    w = load_checkpoint(ckpt)
    for side in (300, 600):
        d0 = synthetic_dense0(side)
        imgs = synthetic_suite(8, side)
        orc = RoomNetOracle(im_side=side, dtype=np.float32, weights=w, dense0_kernel=d0, conv_backend="torch")
        ref = orc.forward(orc.normalise(imgs))
        h = _capi.Handle(im_side=side, precision="fp16")
        h.set_dense0(d0)
        h.load_tf_checkpoint(ckpt)
        t, p, l = h.infer_u8_bgr(imgs, want_logits=True)
        B = 512
        d_in = torch.from_numpy(np.ascontiguousarray(imgs[np.arange(B) % 8])).cuda()
        d_top1 = torch.empty(B, dtype=torch.int64, device="cuda")
        d_probs = torch.empty(B, 6, device="cuda")
        stream = torch.cuda.Stream()
        run = lambda: h.infer_u8_bgr_device(d_in.data_ptr(), B, d_top1.data_ptr(), d_probs.data_ptr(), None,
                                            stream.cuda_stream)
        run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(5):
            run()
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        flop = {300: 8469477408, 600: 36466722336}[side]
        print(json.dumps({"config": 4, "im_side": side, "batch": B, "images_per_s_device": B / ms * 1e3,
                          "tflops": B / ms * 1e3 * flop / 1e12, "top1_equal": bool(np.array_equal(t, ref["argmax"])),
                          "max_abs_dlogit": float(np.abs(l - ref["logits"]).max())}), flush=True)
        h.close()

if "5" in which:
    h = _capi.Handle(precision="fp16", max_batch=1)
    h.load_tf_checkpoint(ckpt)
    x = torch.from_numpy(((synthetic_suite(1).astype(np.float32) - 127.5) / 127.5)).pin_memory()
    pr = torch.empty(1, 6).pin_memory()
    for _ in range(100):
        h.infer_raw("rn_infer_f32_rgb", x.data_ptr(), 1, None, pr.data_ptr(), None)
    h.reset_stats()
    for _ in range(1000):
        h.infer_raw("rn_infer_f32_rgb", x.data_ptr(), 1, None, pr.data_ptr(), None)
    print(json.dumps({"config": 5, "path": "rn_infer_f32_rgb batch 1 (pinned host buffers, incl. H2D/D2H)", **h.stats()}),
          flush=True)
