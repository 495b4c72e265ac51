"""The C-ABI boundary: every symbol include/roomnet.h declares is exported, argument checking, and the
"no CPU fallback" rule (runs without a GPU)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "roomnet.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rn_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(capi):
    lib = ctypes.CDLL(capi.LIB_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 18
    for name in declared:
        assert hasattr(lib, name), "libroomnet.so does not export %s" % name
    assert set(capi.EXPORTED) == set(declared)


def test_jni_library_exports_entry_points():
    lib = ctypes.CDLL(os.path.join(ROOT, "roomnet_b200", "libroomnet_jni.so"))
    for fn in ("create", "run", "runArgb", "runYuv", "close", "stats"):
        assert hasattr(lib, "Java_org_tensorflow_lite_examples_classification_tflite_RoomNetNative_" + fn)


def test_version_and_config_validation(capi):
    assert "sm_100a" in capi.version()
    with pytest.raises(capi.RoomNetError) as e:
        capi.Handle(devices=(), num_classes=0)
    assert e.value.code == capi.RN_ERR_INVALID_ARG
    with pytest.raises(capi.RoomNetError) as e:
        capi.Handle(devices=(), im_side=20)  # too small for the layer stack
    assert e.value.code == capi.RN_ERR_INVALID_ARG and "too small" in str(e.value)
    with pytest.raises(capi.RoomNetError) as e:
        capi.Handle(devices=(), precision=7)
    assert e.value.code == capi.RN_ERR_INVALID_ARG
    cfg = capi.RnConfig()
    cfg.abi_version = 99
    out = ctypes.c_void_p()
    assert capi.lib.rn_create(ctypes.byref(cfg), ctypes.byref(out)) == capi.RN_ERR_INVALID_ARG
    assert b"abi_version" in capi.lib.rn_last_error(None)
    assert capi.lib.rn_create(None, None) == capi.RN_ERR_INVALID_ARG


def test_no_cpu_fallback(capi, ckpt_prefix):
    """Host-only handles load and fold but every inference entry point refuses to run."""
    h = capi.Handle(devices=())
    x = np.zeros((1, 224, 224, 3), np.uint8)
    with pytest.raises(capi.RoomNetError) as e:
        h.infer_u8_bgr(x)
    assert e.value.code == capi.RN_ERR_CUDA and "no CPU inference path" in str(e.value)
    h.load_tf_checkpoint(ckpt_prefix)
    for fn in (h.infer_u8_bgr, h.infer_u8_rgb):
        with pytest.raises(capi.RoomNetError) as e:
            fn(x)
        assert e.value.code == capi.RN_ERR_CUDA
    with pytest.raises(capi.RoomNetError) as e:
        h.infer_f32_rgb(x.astype(np.float32))
    assert e.value.code == capi.RN_ERR_CUDA
    with pytest.raises(capi.RoomNetError):
        h.debug_activation(0)
    import torch
    if not torch.cuda.is_available():
        with pytest.raises(capi.RoomNetError) as e:
            capi.Handle(devices=(0,))
        assert e.value.code == capi.RN_ERR_CUDA


def test_flat_len_matches_reference_shapes(capi):
    # reference network.py:231-232
    assert [capi.Handle(devices=(), im_side=s).flat_len for s in (224, 300, 600)] == [64, 256, 3136]


@pytest.mark.parametrize("h,w", [(224, 224), (480, 640), (640, 480), (641, 480), (480, 641), (3, 1000), (1000, 3)])
def test_center_crop_rect_matches_reference_quirk(capi, h, w):
    from oracle.roomnet_oracle import RoomNetOracle
    img = np.arange(h * w, dtype=np.int64).reshape(h, w, 1)
    want = RoomNetOracle.center_crop(img)
    y0, x0, side = capi.center_crop_rect(h, w)
    got = img[y0:y0 + side, x0:x0 + side]
    assert got.shape == want.shape and np.array_equal(got, want)


def test_null_handle_calls_are_rejected(capi):
    lib = capi.lib
    assert lib.rn_load_tf_checkpoint(None, b"x") == capi.RN_ERR_INVALID_ARG
    assert lib.rn_infer_u8_bgr(None, None, 1, None, None, None) == capi.RN_ERR_INVALID_ARG
    assert lib.rn_flat_len(None) == -1
    assert lib.rn_destroy(None) == capi.RN_OK
    h = capi.Handle(devices=())
    assert lib.rn_load_tf_checkpoint(h._h, None) == capi.RN_ERR_INVALID_ARG
    assert lib.rn_infer_u8_bgr(h._h, None, 1, None, None, None) == capi.RN_ERR_INVALID_ARG
    assert lib.rn_infer_u8_bgr(h._h, ctypes.c_void_p(16), -1, None, None, None) == capi.RN_ERR_INVALID_ARG
