"""Decodes 31 JPEG files of every supported kind on the device (Huffman stage included) and compares each with
cv2.imdecode; the workload of the compute-sanitizer runs on the JPEG kernels (profiles/r02_sanitizer.md)."""
import sys, numpy as np, cv2
import os; ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from roomnet_b200 import _capi
from test_jpeg import photo, encode, CASES, with_exif_orientation, with_16bit_quant_tables
h=_capi.Handle(precision="fp16", max_batch=8)
from roomnet_b200.workload import default_checkpoint_prefix
h.load_tf_checkpoint(default_checkpoint_prefix())
files = [encode(photo(hh, w, seed=hh + w), sf, q, extra) for hh, w, sf, q, extra in CASES[::3]]
names = [str(c) for c in CASES[::3]]
files.append(encode(photo(1213, 1777, seed=9), "420", 92)); names.append("big420")
files.append(encode(photo(900, 1400, seed=10), "422", 85, (cv2.IMWRITE_JPEG_RST_INTERVAL, 7))); names.append("rst422")
files.append(cv2.imencode(".jpg", cv2.cvtColor(photo(300, 500), cv2.COLOR_BGR2GRAY))[1].tobytes()); names.append("grey")
for n, data in zip(names, files):
    got,st=h.decode_jpeg(data)
    ref=cv2.imdecode(np.frombuffer(data,np.uint8),cv2.IMREAD_COLOR)
    print(n, len(data), st, None if got is None else np.array_equal(got,ref), h.jpeg_counters(), flush=True)
