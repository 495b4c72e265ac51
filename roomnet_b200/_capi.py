"""ctypes binding of libroomnet.so (include/roomnet.h).

The library is the product; this module only marshals pointers.  There is no
Python/NumPy fallback: if the shared library is missing, importing fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libroomnet.so")

RN_ABI_VERSION = 2
RN_FLAG_LAYERWISE = 1
RN_FLAG_JPEG_HOST_HUFFMAN = 2
RN_MAX_DEVICES = 16

RN_OK, RN_ERR_INVALID_ARG, RN_ERR_IO, RN_ERR_FORMAT, RN_ERR_NOT_LOADED, RN_ERR_CUDA, RN_ERR_INTERNAL = range(7)
RN_PREC_FP32, RN_PREC_FP16, RN_PREC_BF16, RN_PREC_FP32_TC, RN_PREC_BF16X3 = 0, 1, 2, 3, 4
PRECISIONS = {"fp32": RN_PREC_FP32, "fp16": RN_PREC_FP16, "bf16": RN_PREC_BF16, "fp32tc": RN_PREC_FP32_TC, "bf16x3": RN_PREC_BF16X3}


class RoomNetError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libroomnet error %d: %s" % (code, msg))
        self.code = code


class RnConfig(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32),
        ("im_side", C.c_int32),
        ("num_classes", C.c_int32),
        ("precision", C.c_int32),
        ("n_devices", C.c_int32),
        ("devices", C.c_int32 * RN_MAX_DEVICES),
        ("max_batch", C.c_int32),
        ("flags", C.c_int32),
    ]


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "%s not found — build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C roomnet_b200/csrc` (there is no CPU fallback)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64p, f32p = C.c_void_p, C.c_int32, C.POINTER(C.c_int64), C.POINTER(C.c_float)
    sigs = {
        "rn_create": ([C.POINTER(RnConfig), C.POINTER(vp)], C.c_int),
        "rn_destroy": ([vp], C.c_int),
        "rn_load_tf_checkpoint": ([vp, C.c_char_p], C.c_int),
        "rn_load_tensors": ([vp, i32, C.POINTER(C.c_char_p), C.POINTER(f32p), C.POINTER(i64p), C.POINTER(i32)], C.c_int),
        "rn_set_dense0": ([vp, f32p, i32], C.c_int),
        "rn_infer_u8_bgr": ([vp, vp, i32, vp, vp, vp], C.c_int),
        "rn_infer_u8_rgb": ([vp, vp, i32, vp, vp, vp], C.c_int),
        "rn_infer_f32_rgb": ([vp, vp, i32, vp, vp, vp], C.c_int),
        "rn_infer_argb8888": ([vp, vp, i32, vp, vp, vp], C.c_int),
        "rn_infer_u8_bgr_device": ([vp, vp, i32, vp, vp, vp, vp], C.c_int),
        "rn_submit_u8_bgr": ([vp, vp, i32, vp, vp, vp, C.POINTER(C.c_uint64)], C.c_int),
        "rn_wait": ([vp, C.c_uint64], C.c_int),
        "rn_preprocess_u8": ([vp, vp, i32, i32, vp], C.c_int),
        "rn_infer_image_u8_bgr": ([vp, vp, i32, i32, vp, vp, vp], C.c_int),
        "rn_infer_images_u8_bgr": ([vp, C.POINTER(vp), C.POINTER(i32), C.POINTER(i32), i32, vp, vp, vp], C.c_int),
        "rn_infer_jpeg": ([vp, C.POINTER(vp), C.POINTER(C.c_uint64), i32, i32, vp, vp, vp, vp], C.c_int),
        "rn_decode_jpeg_u8_bgr": ([vp, vp, C.c_uint64, vp, C.c_uint64, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)],
                                  C.c_int),
        "rn_jpeg_info": ([vp, C.c_uint64, i64p], C.c_int),
        "rn_get_jpeg_counters": ([vp, i64p, i64p], C.c_int),
        "rn_jpeg_prepare_scan": ([vp, C.c_uint64, vp, C.c_uint64, vp, i64p], C.c_int),
        "rn_jpeg_coefficients": ([vp, C.c_uint64, vp, C.c_uint64], C.c_int),
        "rn_infer_yuv420": ([vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, i32, i32, vp, vp, vp, vp], C.c_int),
        "rn_center_crop_rect": ([i32, i32, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)], C.c_int),
        "rn_flat_len": ([vp], C.c_int),
        "rn_num_kernel_launches": ([vp], C.c_int),
        "rn_get_folded": ([vp, C.c_char_p, f32p, C.c_int64, i64p], C.c_int),
        "rn_debug_activation": ([vp, i32, f32p, C.c_int64, i64p, C.POINTER(i32)], C.c_int),
        "rn_get_stats": ([vp, C.POINTER(C.c_double), C.POINTER(C.c_double), i64p, i64p], C.c_int),
        "rn_reset_stats": ([vp], C.c_int),
        "rn_set_profiling": ([vp, i32], C.c_int),
        "rn_get_profile": ([vp, i32, C.POINTER(i32), vp, C.POINTER(C.c_double), C.POINTER(i32)], C.c_int),
        "rn_last_error": ([vp], C.c_char_p),
        "rn_version": ([], C.c_char_p),
    }
    for name, (argtypes, restype) in sigs.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = restype
    return lib


lib = _load()
JPEG_OK, JPEG_UNSUPPORTED, JPEG_CORRUPT = 0, 1, 2


def jpeg_info(data):
    """Host-only: (status, width, height, components, luma h, luma v, EXIF orientation, coefficient count)."""
    buf = np.frombuffer(data, dtype=np.uint8)
    info = (C.c_int64 * 8)()
    rc = lib.rn_jpeg_info(buf.ctypes.data, buf.size, info)
    if rc != RN_OK:
        raise RoomNetError(rc, "rn_jpeg_info")
    return tuple(int(v) for v in info)


def jpeg_prepare_scan(data):
    """Host-only: (status, unstuffed stream bytes, restart segment of every 128-byte subsequence, (segments, blocks per
    MCU, total blocks)) - what the device Huffman decoder is given for this file."""
    buf = np.frombuffer(data, dtype=np.uint8)
    cap = (buf.size // 128 + 70000) * 128
    stream = np.zeros(cap, np.uint8)
    sub_seg = np.full(cap // 128, -1, np.int32)
    info = (C.c_int64 * 4)()
    st = lib.rn_jpeg_prepare_scan(buf.ctypes.data, buf.size, stream.ctypes.data, cap, sub_seg.ctypes.data, info)
    if st != JPEG_OK:
        return st, None, None, None
    n = int(info[0])
    return st, stream[:n], sub_seg[:n // 128], (int(info[1]), int(info[2]), int(info[3]))


def jpeg_coefficients(data):
    """Host-only: (status, int16 coefficients or None) - the entropy-decoded, still quantised DCT coefficients."""
    info = jpeg_info(data)
    if info[0] != JPEG_OK:
        return info[0], None
    buf = np.frombuffer(data, dtype=np.uint8)
    coefs = np.empty((info[7],), np.int16)
    st = lib.rn_jpeg_coefficients(buf.ctypes.data, buf.size, coefs.ctypes.data, coefs.size)
    return st, (coefs if st == JPEG_OK else None)
EXPORTED = ["rn_create", "rn_destroy", "rn_load_tf_checkpoint", "rn_load_tensors", "rn_set_dense0",
            "rn_infer_u8_bgr", "rn_infer_u8_rgb", "rn_infer_f32_rgb", "rn_infer_argb8888", "rn_infer_u8_bgr_device",
            "rn_submit_u8_bgr", "rn_wait",
            "rn_preprocess_u8", "rn_infer_image_u8_bgr", "rn_infer_images_u8_bgr", "rn_infer_jpeg",
            "rn_decode_jpeg_u8_bgr", "rn_get_jpeg_counters", "rn_jpeg_info", "rn_jpeg_prepare_scan", "rn_jpeg_coefficients", "rn_infer_yuv420", "rn_center_crop_rect", "rn_flat_len", "rn_num_kernel_launches", "rn_get_folded",
            "rn_debug_activation", "rn_get_stats", "rn_reset_stats", "rn_set_profiling", "rn_get_profile",
            "rn_last_error", "rn_version"]


class Handle:
    """Thin RAII wrapper over rn_handle."""

    def __init__(self, im_side=224, num_classes=6, precision="fp16", devices=(0,), max_batch=0, layerwise=False,
                 jpeg_host_huffman=False):
        cfg = RnConfig()
        cfg.abi_version = RN_ABI_VERSION
        cfg.im_side = im_side
        cfg.num_classes = num_classes
        cfg.precision = PRECISIONS[precision] if isinstance(precision, str) else int(precision)
        devices = list(devices)
        cfg.n_devices = len(devices)
        for i, d in enumerate(devices):
            cfg.devices[i] = d
        cfg.max_batch = max_batch
        cfg.flags = (RN_FLAG_LAYERWISE if layerwise else 0) | (RN_FLAG_JPEG_HOST_HUFFMAN if jpeg_host_huffman else 0)
        self.im_side, self.num_classes = im_side, num_classes
        self._h = C.c_void_p()
        rc = lib.rn_create(C.byref(cfg), C.byref(self._h))
        if rc != RN_OK:
            raise RoomNetError(rc, lib.rn_last_error(None).decode())

    def _check(self, rc):
        if rc != RN_OK:
            raise RoomNetError(rc, lib.rn_last_error(self._h).decode())

    def close(self):
        if getattr(self, "_h", None) and self._h.value and lib is not None:  # lib is None at interpreter shutdown
            lib.rn_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close

    def load_tf_checkpoint(self, prefix):
        self._check(lib.rn_load_tf_checkpoint(self._h, os.fsencode(prefix)))

    def load_tensors(self, tensors: dict):
        names = list(tensors)
        arrs = [np.ascontiguousarray(tensors[k], dtype=np.float32) for k in names]
        shapes = [np.asarray(a.shape, dtype=np.int64) for a in arrs]
        n = len(names)
        c_names = (C.c_char_p * n)(*[k.encode() for k in names])
        c_data = (C.POINTER(C.c_float) * n)(*[a.ctypes.data_as(C.POINTER(C.c_float)) for a in arrs])
        c_shapes = (C.POINTER(C.c_int64) * n)(*[s.ctypes.data_as(C.POINTER(C.c_int64)) for s in shapes])
        c_ranks = (C.c_int32 * n)(*[a.ndim for a in arrs])
        self._check(lib.rn_load_tensors(self._h, n, c_names, c_data, c_shapes, c_ranks))

    def set_dense0(self, kernel):
        k = np.ascontiguousarray(kernel, dtype=np.float32)
        self._check(lib.rn_set_dense0(self._h, k.ctypes.data_as(C.POINTER(C.c_float)), k.shape[0]))

    def _infer(self, fn, x, dtype, want_logits):
        x = np.asarray(x)
        if x.ndim != 4 or x.shape[1:] != (self.im_side, self.im_side, 3):
            # TF raises InvalidArgumentError for a feed that does not match the placeholder shape
            raise RoomNetError(RN_ERR_INVALID_ARG, "input must be [n,%d,%d,3], got %s"
                               % (self.im_side, self.im_side, x.shape))
        x = np.ascontiguousarray(x, dtype=dtype)
        n = x.shape[0]
        top1 = np.empty((n,), np.int64)
        probs = np.empty((n, self.num_classes), np.float32)
        logits = np.empty((n, self.num_classes), np.float32) if want_logits else None
        self._check(fn(self._h, x.ctypes.data, n, top1.ctypes.data, probs.ctypes.data,
                       logits.ctypes.data if want_logits else None))
        return (top1, probs, logits) if want_logits else (top1, probs)

    def infer_u8_bgr(self, x, want_logits=False):
        return self._infer(lib.rn_infer_u8_bgr, x, np.uint8, want_logits)

    def infer_u8_rgb(self, x, want_logits=False):
        return self._infer(lib.rn_infer_u8_rgb, x, np.uint8, want_logits)

    def infer_f32_rgb(self, x, want_logits=False):
        return self._infer(lib.rn_infer_f32_rgb, x, np.float32, want_logits)

    def infer_argb8888(self, pixels, want_logits=False):
        """[n, S, S] int32 0xAARRGGBB pixels (android.graphics.Bitmap.getPixels)."""
        x = np.asarray(pixels)
        if x.ndim != 3 or x.shape[1:] != (self.im_side, self.im_side):
            raise RoomNetError(RN_ERR_INVALID_ARG, "pixels must be [n,%d,%d] int32, got %s"
                               % (self.im_side, self.im_side, x.shape))
        x = np.ascontiguousarray(x, dtype=np.int32)
        n = x.shape[0]
        top1 = np.empty((n,), np.int64)
        probs = np.empty((n, self.num_classes), np.float32)
        logits = np.empty((n, self.num_classes), np.float32) if want_logits else None
        self._check(lib.rn_infer_argb8888(self._h, x.ctypes.data, n, top1.ctypes.data, probs.ctypes.data,
                                          logits.ctypes.data if want_logits else None))
        return (top1, probs, logits) if want_logits else (top1, probs)

    def preprocess_u8(self, img):
        """center_crop + cv2.resize(img, (S, S)) on the device (bit-identical to cv2 for uint8)."""
        img = np.ascontiguousarray(img, dtype=np.uint8)
        if img.ndim != 3 or img.shape[2] != 3:
            raise RoomNetError(RN_ERR_INVALID_ARG, "image must be [H, W, 3] uint8, got %s" % (img.shape,))
        out = np.empty((self.im_side, self.im_side, 3), np.uint8)
        self._check(lib.rn_preprocess_u8(self._h, img.ctypes.data, img.shape[0], img.shape[1], out.ctypes.data))
        return out

    def infer_image_u8_bgr(self, img, want_logits=False):
        img = np.ascontiguousarray(img, dtype=np.uint8)
        if img.ndim != 3 or img.shape[2] != 3:
            raise RoomNetError(RN_ERR_INVALID_ARG, "image must be [H, W, 3] uint8, got %s" % (img.shape,))
        top1 = np.empty((1,), np.int64)
        probs = np.empty((1, self.num_classes), np.float32)
        logits = np.empty((1, self.num_classes), np.float32)
        self._check(lib.rn_infer_image_u8_bgr(self._h, img.ctypes.data, img.shape[0], img.shape[1], top1.ctypes.data,
                                              probs.ctypes.data, logits.ctypes.data))
        return (top1, probs, logits) if want_logits else (top1, probs)

    def infer_images_u8_bgr(self, imgs, want_logits=False):
        """List of BGR uint8 images of any size: crop + resize + forward pass on the device, one call."""
        arrs = []
        for img in imgs:
            a = np.ascontiguousarray(img, dtype=np.uint8)
            if a.ndim != 3 or a.shape[2] != 3:
                raise RoomNetError(RN_ERR_INVALID_ARG, "image must be [H, W, 3] uint8, got %s" % (a.shape,))
            arrs.append(a)
        n = len(arrs)
        ptrs = (C.c_void_p * n)(*[a.ctypes.data for a in arrs])
        hs = (C.c_int32 * n)(*[a.shape[0] for a in arrs])
        ws = (C.c_int32 * n)(*[a.shape[1] for a in arrs])
        top1 = np.empty((n,), np.int64)
        probs = np.empty((n, self.num_classes), np.float32)
        logits = np.empty((n, self.num_classes), np.float32)
        self._check(lib.rn_infer_images_u8_bgr(self._h, ptrs, hs, ws, n, top1.ctypes.data, probs.ctypes.data,
                                               logits.ctypes.data))
        return (top1, probs, logits) if want_logits else (top1, probs)

    def infer_jpeg(self, files, threads=0, want_logits=False):
        """Encoded files (bytes objects): entropy decoding on host threads, the rest of the decoder + crop + resize +
        forward pass on the device.  Returns (top1, probs[, logits], status); rows whose status is not JPEG_OK are
        not filled in (decode those files on the host and use infer_images_u8_bgr)."""
        n = len(files)
        bufs = [np.frombuffer(f, dtype=np.uint8) for f in files]
        ptrs = (C.c_void_p * n)(*[b.ctypes.data for b in bufs])
        sizes = (C.c_uint64 * n)(*[b.size for b in bufs])
        top1 = np.full((n,), -1, np.int64)
        probs = np.zeros((n, self.num_classes), np.float32)
        logits = np.zeros((n, self.num_classes), np.float32)
        status = np.full((n,), JPEG_UNSUPPORTED, np.int32)
        self._check(lib.rn_infer_jpeg(self._h, ptrs, sizes, n, threads, top1.ctypes.data, probs.ctypes.data,
                                      logits.ctypes.data, status.ctypes.data))
        return (top1, probs, logits, status) if want_logits else (top1, probs, status)

    def jpeg_counters(self):
        """(files Huffman-decoded on the device, files Huffman-decoded on host threads) so far."""
        d, c = C.c_int64(), C.c_int64()
        self._check(lib.rn_get_jpeg_counters(self._h, C.byref(d), C.byref(c)))
        return d.value, c.value

    def decode_jpeg(self, data):
        """cv2.imdecode(data, cv2.IMREAD_COLOR) for a baseline JPEG, second half of the decoder on the device.
        Returns (image or None, status)."""
        buf = np.frombuffer(data, dtype=np.uint8)
        hh, ww, st = C.c_int32(), C.c_int32(), C.c_int32()
        self._check(lib.rn_decode_jpeg_u8_bgr(self._h, buf.ctypes.data, buf.size, None, 0, C.byref(hh), C.byref(ww),
                                              C.byref(st)))
        if st.value != JPEG_OK:
            return None, st.value
        out = np.empty((hh.value, ww.value, 3), np.uint8)
        self._check(lib.rn_decode_jpeg_u8_bgr(self._h, buf.ctypes.data, buf.size, out.ctypes.data, out.nbytes,
                                              C.byref(hh), C.byref(ww), C.byref(st)))
        return (out if st.value == JPEG_OK else None), st.value

    def infer_yuv420(self, y, u, v, width, height, y_row_stride, uv_row_stride, uv_pixel_stride, rotation=0):
        """One YUV_420_888 camera frame (planes as uint8 arrays) -> (top1, probs, logits, rgb the network saw)."""
        y, u, v = (np.ascontiguousarray(a, dtype=np.uint8).ravel() for a in (y, u, v))
        top1 = np.empty((1,), np.int64)
        probs = np.empty((1, self.num_classes), np.float32)
        logits = np.empty((1, self.num_classes), np.float32)
        rgb = np.empty((self.im_side, self.im_side, 3), np.uint8)
        self._check(lib.rn_infer_yuv420(self._h, y.ctypes.data, u.ctypes.data, v.ctypes.data, y.size, u.size, v.size,
                                        width, height, y_row_stride, uv_row_stride, uv_pixel_stride, rotation,
                                        top1.ctypes.data, probs.ctypes.data, logits.ctypes.data, rgb.ctypes.data))
        return top1, probs, logits, rgb

    def infer_raw(self, fn_name, in_ptr, n, top1_ptr, probs_ptr, logits_ptr):
        """Pointer-level call (pinned host buffers owned by the caller)."""
        self._check(getattr(lib, fn_name)(self._h, in_ptr, n, top1_ptr, probs_ptr, logits_ptr))

    def submit_raw(self, in_ptr, n, top1_ptr, probs_ptr, logits_ptr):
        """rn_submit_u8_bgr on caller-owned (ideally pinned) buffers; returns the ticket for wait()."""
        ticket = C.c_uint64()
        self._check(lib.rn_submit_u8_bgr(self._h, in_ptr, n, top1_ptr, probs_ptr, logits_ptr, C.byref(ticket)))
        return ticket.value

    def wait(self, ticket=0):
        """Block until every call up to `ticket` (0 = all) has delivered its results."""
        self._check(lib.rn_wait(self._h, ticket))

    def infer_u8_bgr_device(self, d_in, n, d_top1, d_probs, d_logits, stream=None):
        self._check(lib.rn_infer_u8_bgr_device(self._h, d_in, n, d_top1, d_probs, d_logits, stream))

    @property
    def flat_len(self):
        return lib.rn_flat_len(self._h)

    @property
    def kernel_launches(self):
        return lib.rn_num_kernel_launches(self._h)

    def get_folded(self, name):
        size = C.c_int64()
        self._check(lib.rn_get_folded(self._h, name.encode(), None, 0, C.byref(size)))
        out = np.empty((size.value,), np.float32)
        self._check(lib.rn_get_folded(self._h, name.encode(), out.ctypes.data_as(C.POINTER(C.c_float)),
                                      size.value, C.byref(size)))
        return out

    def debug_activation(self, layer):
        size = C.c_int64()
        dims = (C.c_int32 * 4)()
        self._check(lib.rn_debug_activation(self._h, layer, None, 0, C.byref(size), dims))
        out = np.empty((size.value,), np.float32)
        self._check(lib.rn_debug_activation(self._h, layer, out.ctypes.data_as(C.POINTER(C.c_float)),
                                            size.value, C.byref(size), dims))
        return out.reshape(tuple(dims))

    def stats(self):
        p50, p99 = C.c_double(), C.c_double()
        calls, images = C.c_int64(), C.c_int64()
        self._check(lib.rn_get_stats(self._h, C.byref(p50), C.byref(p99), C.byref(calls), C.byref(images)))
        return dict(p50_ms=p50.value, p99_ms=p99.value, calls=calls.value, images=images.value)

    def set_profiling(self, on):
        self._check(lib.rn_set_profiling(self._h, 1 if on else 0))

    def get_profile(self):
        cap = 64
        names = (C.c_char * 32 * cap)()
        ms = (C.c_double * cap)()
        launches = (C.c_int32 * cap)()
        count = C.c_int32()
        self._check(lib.rn_get_profile(self._h, cap, C.byref(count), C.cast(names, C.c_void_p), ms, launches))
        return [dict(name=names[i].value.decode(), ms=ms[i], launches=launches[i]) for i in range(min(cap, count.value))]

    def reset_stats(self):
        self._check(lib.rn_reset_stats(self._h))


def center_crop_rect(h, w):
    y0, x0, side = C.c_int32(), C.c_int32(), C.c_int32()
    rc = lib.rn_center_crop_rect(h, w, C.byref(y0), C.byref(x0), C.byref(side))
    if rc != RN_OK:
        raise RoomNetError(rc, "bad image size")
    return y0.value, x0.value, side.value


def version():
    return lib.rn_version().decode()
