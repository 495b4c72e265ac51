"""Front end of infer_optimized (reference network.py:137-152): centre crop + cv2.resize.

CPU: the integer restatement of OpenCV's uint8 INTER_LINEAR path (oracle/cv_resize.py) is pinned bit-exactly to the
installed cv2.  GPU: rn_preprocess_u8 / rn_infer_image_u8_bgr are bit-identical to the cv2 host path."""
import numpy as np
import pytest

SHAPES = [(480, 640), (641, 480), (480, 641), (300, 300), (448, 448), (100, 100), (225, 225), (223, 224), (600, 600),
          (50, 400), (449, 447), (37, 91), (224, 224), (1080, 1920), (3, 5)]


def _images(seed=0):
    rng = np.random.default_rng(seed)
    return [rng.integers(0, 256, s + (3,), dtype=np.uint8) for s in SHAPES]


@pytest.mark.parametrize("side", [224, 300])
def test_oracle_resize_is_bit_exact_vs_cv2(side):
    import cv2
    from oracle.cv_resize import preprocess
    from oracle.roomnet_oracle import RoomNetOracle
    for img in _images():
        crop = RoomNetOracle.center_crop(img)
        want = crop if crop.shape[:2] == (side, side) else cv2.resize(crop, (side, side))
        assert np.array_equal(preprocess(img, side), want), img.shape


def test_oracle_resize_random_shapes_vs_cv2():
    import cv2
    from oracle.cv_resize import resize_linear_u8
    rng = np.random.default_rng(3)
    for _ in range(25):
        s = int(rng.integers(3, 900))
        img = rng.integers(0, 256, (s, s, 3), dtype=np.uint8)
        assert np.array_equal(resize_linear_u8(img, 224, 224), cv2.resize(img, (224, 224))), s


@pytest.mark.gpu
@pytest.mark.parametrize("side", [224, 300])
def test_gpu_preprocess_is_bit_exact_vs_cv2(capi, side):
    from oracle.cv_resize import preprocess
    h = capi.Handle(im_side=side, precision="fp32", max_batch=2)
    for img in _images(1):
        got = h.preprocess_u8(img)
        assert got.shape == (side, side, 3)
        assert np.array_equal(got, preprocess(img, side)), img.shape


@pytest.mark.gpu
def test_gpu_infer_image_equals_host_preprocess_path(ckpt_prefix):
    from roomnet_b200 import RoomNet
    host = RoomNet(num_classes=6, im_side=224, compute_bn_mean_var=False, optimized_inference=True)
    dev = RoomNet(num_classes=6, im_side=224, compute_bn_mean_var=False, optimized_inference=True, gpu_preprocess=True)
    host.load(ckpt_prefix)
    dev.load(ckpt_prefix)
    for img in _images(2)[:8]:
        i0, c0 = host.infer_optimized(img)
        i1, c1 = dev.infer_optimized(img)
        assert i0.shape == i1.shape == (1,) and c1.shape == (1, 6)
        assert np.array_equal(i0, i1) and np.array_equal(c0, c1), img.shape
    b0 = host.infer_optimized_batch(_images(2)[:5])
    b1 = dev.infer_optimized_batch(_images(2)[:5])
    assert np.array_equal(b0[0], b1[0]) and np.array_equal(b0[1], b1[1])
