"""Row n1 measurement: a directory's worth of JPEG photographs -> labels, decoded by cv2 on host threads (what
classify_im_dir did before) against rn_infer_jpeg (entropy decoding on host threads, the rest of the decoder on the
device).  Run under gpurun; prints one JSON line."""
import json
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

import cv2
import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from roomnet_b200 import _capi  # noqa: E402
from roomnet_b200.workload import default_checkpoint_prefix  # noqa: E402


def photo(h, w, seed):
    rng = np.random.default_rng(seed)
    small = rng.integers(0, 256, (h // 16 + 2, w // 16 + 2, 3), dtype=np.uint8)
    img = cv2.resize(small, (w, h), interpolation=cv2.INTER_CUBIC).astype(np.int16)
    img += rng.integers(-10, 11, img.shape, dtype=np.int16)  # sensor-noise-like texture: realistic entropy
    return np.clip(img, 0, 255).astype(np.uint8)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 96
    sizes = [(3000, 4000), (2448, 3264), (1080, 1920), (1536, 2048)]
    files = []
    for i in range(n):
        h, w = sizes[i % len(sizes)]
        ok, enc = cv2.imencode(".jpg", photo(h, w, i), [cv2.IMWRITE_JPEG_QUALITY, 90])
        files.append(enc.tobytes())
    mpix = sum(sizes[i % len(sizes)][0] * sizes[i % len(sizes)][1] for i in range(n)) / 1e6
    mbytes = sum(len(f) for f in files) / 1e6
    h = _capi.Handle(precision="fp16", max_batch=64)
    h.load_tf_checkpoint(default_checkpoint_prefix())
    hh = _capi.Handle(precision="fp16", max_batch=64, jpeg_host_huffman=True)
    hh.load_tf_checkpoint(default_checkpoint_prefix())
    threads = min(16, os.cpu_count() or 1)
    cv2.setNumThreads(1)

    def host_path():
        with ThreadPoolExecutor(max_workers=threads) as pool:
            out = []
            for lo in range(0, n, 64):
                ims = list(pool.map(lambda f: cv2.imdecode(np.frombuffer(f, np.uint8), cv2.IMREAD_COLOR), files[lo:lo + 64]))
                out.append(h.infer_images_u8_bgr(ims, want_logits=True)[2])
        return np.concatenate(out)

    def device_path():
        return h.infer_jpeg(files, threads=threads, want_logits=True)

    def host_huffman_path():
        return hh.infer_jpeg(files, threads=threads, want_logits=True)

    ref = host_path()
    t1, p1, l1, st = device_path()
    assert (st == 0).all() and np.array_equal(l1, ref), "device decode must be bit-identical to cv2's"
    assert h.jpeg_counters() == (n, 0), "every file of this set is a single scan: Huffman decoding on the device"
    assert np.array_equal(host_huffman_path()[2], ref) and hh.jpeg_counters() == (0, n)
    res = {}
    for name, fn in (("cv2_host_decode", host_path), ("device_decode", device_path),
                     ("device_decode_host_huffman", host_huffman_path)):
        ts = []
        for _ in range(3):
            t0 = time.perf_counter()
            fn()
            ts.append(time.perf_counter() - t0)
        res[name] = {"s": min(ts), "images_per_s": n / min(ts), "mpix_per_s": mpix / min(ts)}
    # the host half alone on the same number of threads: the floor of the device path on this machine
    with ThreadPoolExecutor(max_workers=threads) as pool:
        list(pool.map(_capi.jpeg_coefficients, files[:threads]))
        t0 = time.perf_counter()
        list(pool.map(_capi.jpeg_coefficients, files))
        res["host_entropy_decode_all_threads"] = {"s": time.perf_counter() - t0}
    # host entropy decoding alone (one thread), for the split of the device path
    t0 = time.perf_counter()
    for f in files[:16]:
        _capi.jpeg_coefficients(f)
    ent = (time.perf_counter() - t0) / 16
    t0 = time.perf_counter()
    for f in files[:16]:
        cv2.imdecode(np.frombuffer(f, np.uint8), cv2.IMREAD_COLOR)
    full = (time.perf_counter() - t0) / 16
    print(json.dumps({"files": n, "megapixels": round(mpix, 1), "encoded_MB": round(mbytes, 1), "host_threads": threads,
                      "cpu_count": os.cpu_count(), "bit_identical_logits": True, **res,
                      "one_thread_ms_per_file": {"entropy_decode_only": round(ent * 1e3, 2),
                                                 "cv2_imdecode": round(full * 1e3, 2)},
                      "speedup": res["device_decode"]["images_per_s"] / res["cv2_host_decode"]["images_per_s"]}))


if __name__ == "__main__":
    main()
