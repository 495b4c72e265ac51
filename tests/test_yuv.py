"""Camera front end of the mobile module (ImageUtils.java:100-151, :168-225; ClassifierActivity.java:89-106) on the
device: rn_infer_yuv420 against the integer restatement in oracle/yuv_front.py."""
import numpy as np
import pytest

from oracle import yuv_front


def _frame(rng, width, height, pixel_stride, pad):
    """A YUV_420_888 frame as android.media.Image hands it out: planar (pixel stride 1) or semi-planar with the U and
    V buffers being two views of one interleaved array (pixel stride 2), row strides with padding."""
    ys = width + pad
    y = rng.integers(0, 256, ys * height, dtype=np.uint8)
    ch, cw = (height + 1) // 2, (width + 1) // 2
    if pixel_stride == 1:
        uvs = cw + pad
        u = rng.integers(0, 256, uvs * ch, dtype=np.uint8)
        v = rng.integers(0, 256, uvs * ch, dtype=np.uint8)
    else:
        uvs = 2 * cw + pad
        inter = rng.integers(0, 256, uvs * ch + 1, dtype=np.uint8)
        v, u = inter[:-1], inter[1:]  # NV21: V first
    return y, u, v, ys, uvs


def test_yuv2rgb_known_answers():
    # ImageUtils.YUV2RGB by hand: video-range black, white, and the four chroma extremes
    r, g, b = yuv_front.yuv2rgb(np.array([16, 235, 81, 145, 41, 0]), np.array([128, 128, 90, 54, 240, 0]),
                                np.array([128, 128, 240, 34, 110, 255]))
    assert r.tolist() == [0, 254, 254, 0, 0, 202] and g.tolist() == [0, 254, 0, 255, 0, 0]
    assert b.tolist() == [0, 254, 0, 0, 255, 0]


def test_frame_to_crop_geometry():
    # rotation 0 keeps the top-left corner (no centring in getTransformationMatrix without a rotation)
    img = np.arange(480 * 640 * 3, dtype=np.int64).reshape(480, 640, 3)
    c0 = yuv_front.frame_to_crop(img, 224, 0)
    assert (c0[0, 0] == img[1, 1]).all()  # scale 224/480: centre of crop pixel 0 -> frame 1.07
    # rotation 90 (forward (x, y) -> (-y, x)): the crop centre is the frame centre, one crop pixel = 2.14 frame pixels
    c90 = yuv_front.frame_to_crop(img, 224, 90)
    assert (c90[112, 112] == img[238, 321]).all() and (c90[112, 113] == img[236, 321]).all()
    c180 = yuv_front.frame_to_crop(img, 224, 180)
    assert (c180[111, 111] == img[241, 321]).all()


@pytest.mark.gpu
def test_gpu_yuv_front_end_is_bit_exact_and_feeds_the_same_network(capi, ckpt_prefix):
    h = capi.Handle(im_side=224, precision="fp16", max_batch=1)
    h.load_tf_checkpoint(ckpt_prefix)
    rng = np.random.default_rng(4)
    cases = [(640, 480, 2, 0, 90), (640, 480, 1, 0, 0), (641, 481, 2, 31, 270), (320, 240, 1, 16, 180),
             (224, 224, 2, 0, 0), (1280, 720, 2, 0, 90), (300, 500, 1, 3, 90)]
    for width, height, ps, pad, rot in cases:
        y, u, v, ys, uvs = _frame(rng, width, height, ps, pad)
        t, p, l, rgb = h.infer_yuv420(y, u, v, width, height, ys, uvs, ps, rot)
        want = yuv_front.camera_front(y, u, v, width, height, ys, uvs, ps, 224, rot)
        assert np.array_equal(rgb, want), (width, height, ps, pad, rot)
        t2, p2, l2 = h.infer_u8_rgb(rgb[None], want_logits=True)
        assert np.array_equal(t, t2) and np.array_equal(p, p2) and np.array_equal(l, l2)
    y, u, v, ys, uvs = _frame(rng, 64, 48, 2, 0)
    with pytest.raises(capi.RoomNetError):
        h.infer_yuv420(y, u, v, 64, 48, ys, uvs, 2, 45)       # not a multiple of 90
    with pytest.raises(capi.RoomNetError):
        h.infer_yuv420(y[:100], u, v, 64, 48, ys, uvs, 2, 0)  # plane too small for the geometry
