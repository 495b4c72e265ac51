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


@pytest.mark.gpu
def test_gpu_batched_images_call_is_bit_identical_to_the_per_image_path(capi, ckpt_prefix):
    """rn_infer_images_u8_bgr: one crop+resize launch per micro-batch for a list of photos of different sizes, chunked
    by max_batch, against rn_infer_image_u8_bgr per image and against cv2 on the host + rn_infer_u8_bgr."""
    from oracle.cv_resize import preprocess
    h = capi.Handle(im_side=224, precision="fp16", max_batch=8)
    h.load_tf_checkpoint(ckpt_prefix)
    rng = np.random.default_rng(11)
    imgs = _images(5) + _images(6)[:7]
    imgs += [rng.integers(0, 256, (int(rng.integers(2, 700)), int(rng.integers(2, 700)), 3), dtype=np.uint8)
             for _ in range(9)]
    t, p, l = h.infer_images_u8_bgr(imgs, want_logits=True)
    assert t.shape == (len(imgs),) and p.shape == (len(imgs), 6)
    host = np.stack([preprocess(im, 224) for im in imgs])
    t2, p2, l2 = h.infer_u8_bgr(host, want_logits=True)
    assert np.array_equal(t, t2) and np.array_equal(p, p2) and np.array_equal(l, l2)
    for i in (0, 3, len(imgs) - 1):
        t1, p1, l1 = h.infer_image_u8_bgr(imgs[i], want_logits=True)
        assert np.array_equal(t1[0], t[i]) and np.array_equal(l1[0], l[i])
    with pytest.raises(capi.RoomNetError):
        h.infer_images_u8_bgr([np.zeros((1, 5, 3), np.uint8)])  # degenerate size
    assert h.infer_images_u8_bgr([])[0].shape == (0,)


@pytest.mark.gpu
def test_gpu_classify_im_dir_on_real_files(ckpt_prefix, tmp_path):
    """The reference's entry point (infer.py:65-100) end to end on the device: image files on disk ->
    <dir>_classified/<Label>/<file> + results table, labels equal to the CPU oracle's on the decoded pixels."""
    import os
    import cv2
    from oracle.cv_resize import preprocess
    from oracle.roomnet_oracle import RoomNetOracle, synthetic_suite
    from roomnet_b200 import RoomNet
    from roomnet_b200.infer import CLASS_LABELS, classify_im_dir
    imgs_dir = tmp_path / "photos"
    imgs_dir.mkdir()
    suite = synthetic_suite(64)
    rng = np.random.default_rng(2)
    names = []
    for i in range(20):
        base = suite[(5 * i + 1) % 64]
        hh, ww = int(rng.integers(230, 520)), int(rng.integers(230, 520))
        img = cv2.resize(base, (ww, hh), interpolation=cv2.INTER_CUBIC)
        name = "im_%02d.%s" % (i, "png" if i % 3 else "jpg")
        cv2.imwrite(str(imgs_dir / name), img)
        names.append(name)
    oracle = RoomNetOracle(dtype=np.float32).load()
    decoded = [cv2.imread(str(imgs_dir / n)) for n in names]
    ref = oracle.forward(oracle.normalise(np.stack([preprocess(d, 224) for d in decoded])))
    for gpu_pre in (True, False):
        nn = RoomNet(num_classes=6, im_side=224, compute_bn_mean_var=False, optimized_inference=True,
                     gpu_preprocess=gpu_pre)
        nn.load(ckpt_prefix)
        xls = classify_im_dir(nn, str(imgs_dir))
        nn.close()
        out_dir = str(imgs_dir) + "_classified"
        assert xls == out_dir + "_results.xls" and os.path.getsize(xls) > 0
        placed = {f: lab for lab in CLASS_LABELS for f in os.listdir(os.path.join(out_dir, lab))}
        assert sorted(placed) == sorted(names)
        margin = np.sort(ref["softmax"], axis=1)
        for n, want, m in zip(names, ref["argmax"], margin[:, -1] - margin[:, -2]):
            if m > 0.02:  # away from ties the label is the oracle's
                assert placed[n] == CLASS_LABELS[int(want)], n
            over = cv2.imread(os.path.join(out_dir, placed[n], n))
            assert over is not None and over.shape == decoded[names.index(n)].shape
        import shutil
        shutil.rmtree(out_dir)
    # overlay=False: the files go to the library as encoded bytes - the JPEGs are decoded on the device, the PNGs by
    # cv2 on the host - and are copied verbatim; same labels as above
    nn = RoomNet(num_classes=6, im_side=224, compute_bn_mean_var=False, optimized_inference=True)
    nn.load(ckpt_prefix)
    classify_im_dir(nn, str(imgs_dir), overlay=False)
    nn.close()
    placed2 = {f: lab for lab in CLASS_LABELS for f in os.listdir(os.path.join(out_dir, lab))}
    assert placed2 == placed
    for n in names:
        assert open(os.path.join(out_dir, placed2[n], n), "rb").read() == open(imgs_dir / n, "rb").read()
