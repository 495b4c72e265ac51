"""JNI shim (reference mobile/.../tflite/Classifier.java entry point) driven by a fake JNIEnv table."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
HARNESS = os.path.join(ROOT, "build", "jni_harness")
JNI_LIB = os.path.join(ROOT, "roomnet_b200", "libroomnet_jni.so")


@pytest.fixture(scope="module")
def harness():
    if not os.path.exists(HARNESS) or os.path.getmtime(HARNESS) < os.path.getmtime(os.path.join(ROOT, "tests", "jni_harness.cpp")):
        os.makedirs(os.path.dirname(HARNESS), exist_ok=True)
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-I", os.path.join(ROOT, "roomnet_b200", "csrc"),
                               os.path.join(ROOT, "tests", "jni_harness.cpp"), "-o", HARNESS, "-ldl"])
    return HARNESS


def test_jni_error_paths(harness, ckpt_prefix):
    out = subprocess.run([harness, JNI_LIB, ckpt_prefix, "errors"], capture_output=True, text=True, timeout=120)
    print(out.stdout, out.stderr)
    assert out.returncode == 0 and "JNI error paths OK" in out.stdout
    assert "java/io/IOException" in out.stdout


@pytest.mark.gpu
def test_jni_run_matches_python_path(harness, ckpt_prefix, capi):
    out = subprocess.run([harness, JNI_LIB, ckpt_prefix, "run", "300"], capture_output=True, text=True, timeout=300)
    print(out.stdout, out.stderr)
    assert out.returncode == 0 and "JNI run OK" in out.stdout
    probs = [list(map(float, l.split()[1:])) for l in out.stdout.splitlines() if l.startswith("probs_f32")][0]
    # same image through the ctypes path
    n = 224 * 224 * 3
    b = ((np.arange(n, dtype=np.uint64) * 2654435761) & 0xFFFFFFFF) >> 24
    img = b.astype(np.uint8).reshape(1, 224, 224, 3)
    h = capi.Handle(precision="fp16", max_batch=1)
    h.load_tf_checkpoint(ckpt_prefix)
    _, p = h.infer_f32_rgb(((img.astype(np.float32) - 127.5) / 127.5))
    assert np.abs(np.array(probs) - p[0]).max() < 1e-5
    # the camera frame of the harness (640x480, NV21 layout, rotation 90) through the ctypes path
    W, H = 640, 480
    yplane = (((np.arange(W * H, dtype=np.uint64) * 2246822519) & 0xFFFFFFFF) >> 24).astype(np.uint8)
    uv = (((np.arange(W * (H // 2) + 1, dtype=np.uint64) * 3266489917) & 0xFFFFFFFF) >> 24).astype(np.uint8)
    py = [list(map(float, l.split()[1:])) for l in out.stdout.splitlines() if l.startswith("probs_yuv")][0]
    _, p2, _, _ = h.infer_yuv420(yplane, uv[1:], uv[:-1], W, H, W, W, 2, 90)
    assert np.abs(np.array(py) - p2[0]).max() < 1e-5
