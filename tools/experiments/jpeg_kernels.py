"""One rn_infer_jpeg call over 32 photographs (for an ncu launch list of the JPEG kernels)."""
import os, sys
import cv2
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from roomnet_b200 import _capi
from roomnet_b200.workload import default_checkpoint_prefix
from bench_jpeg import photo
sizes = [(3000, 4000), (2448, 3264), (1080, 1920), (1536, 2048)]
files = [cv2.imencode(".jpg", photo(*sizes[i % 4], i), [cv2.IMWRITE_JPEG_QUALITY, 90])[1].tobytes() for i in range(32)]
h = _capi.Handle(precision="fp16", max_batch=64)
h.load_tf_checkpoint(default_checkpoint_prefix())
for _ in range(2):
    t, p, st = h.infer_jpeg(files)
print(t[:8], h.jpeg_counters(), sum(len(f) for f in files) / 1e6, "MB")
