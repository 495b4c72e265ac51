"""Host-side mirror of the reference interface (RoomNet class, classify_im_dir) — no GPU needed:
the device call is replaced by a stub so that only the host logic is under test."""
import os
import struct

import cv2
import numpy as np
import pytest

from roomnet_b200 import CLASS_LABELS, RoomNet, classify_im_dir
from roomnet_b200 import infer as infer_mod


def test_reference_constants():
    assert CLASS_LABELS == ['Backyard', 'Bathroom', 'Bedroom', 'Frontyard', 'Kitchen', 'LivingRoom']  # infer.py:22
    assert infer_mod.IMG_SIDE == 224 and infer_mod.INPUT_MODEL_PATH == './final_model/roomnet'


def test_constructor_mirrors_reference_kwargs():
    nn = RoomNet(num_classes=6, im_side=224, compute_bn_mean_var=False, optimized_inference=True)
    assert nn.num_classes == 6 and nn.im_side == 224 and nn.sess is None
    with pytest.raises(NotImplementedError):
        RoomNet(num_classes=6)  # compute_bn_mean_var defaults to True = training-mode BN (network.py:22)
    with pytest.raises(NotImplementedError):
        nn.train_step(None, None)
    with pytest.raises(NotImplementedError):
        nn.save()


def test_center_crop_and_preprocess_equal_reference_semantics():
    from oracle.roomnet_oracle import RoomNetOracle
    nn = RoomNet(num_classes=6, im_side=224, compute_bn_mean_var=False, optimized_inference=True)
    orc = RoomNetOracle(im_side=224)
    rng = np.random.default_rng(0)
    for shape in [(224, 224), (480, 640), (641, 480), (480, 641), (100, 333)]:
        im = rng.integers(0, 256, shape + (3,), dtype=np.uint8)
        assert np.array_equal(nn.center_crop(im), orc.center_crop(im))
        assert np.array_equal(nn.preprocess(im), orc.preprocess(im))
        assert nn.preprocess(im).shape == (224, 224, 3)


def test_infer_before_load_raises():
    from roomnet_b200 import _capi
    nn = RoomNet(num_classes=6, im_side=224, compute_bn_mean_var=False, optimized_inference=True)
    with pytest.raises(_capi.RoomNetError):
        nn.infer(np.zeros((1, 224, 224, 3), np.uint8))


class _StubSession:
    """Stands in for the libroomnet handle: classifies by mean brightness so outputs are predictable."""

    def __init__(self):
        self.batches = []

    def infer_u8_bgr(self, batch, want_logits=False):
        self.batches.append(batch.shape)
        assert batch.dtype == np.uint8 and batch.shape[1:] == (224, 224, 3)
        idx = (batch.reshape(len(batch), -1).mean(axis=1) // 43).astype(np.int64).clip(0, 5)
        probs = np.full((len(batch), 6), 0.02, np.float32)
        probs[np.arange(len(batch)), idx] = 0.9
        return idx, probs

    def infer_jpeg(self, blobs, threads=0, want_logits=False):
        n = len(blobs)  # a session without the device decoder: every file goes back to cv2 on the host
        return np.full(n, -1, np.int64), np.zeros((n, 6), np.float32), np.ones(n, np.int32)


def test_classify_im_dir_outputs(tmp_path):
    imgs_dir = tmp_path / "images"
    imgs_dir.mkdir()
    levels = {"a.png": 10, "b.png": 100, "c.png": 250, "d.jpg": 60}
    for name, v in levels.items():
        cv2.imwrite(str(imgs_dir / name), np.full((300, 400, 3), v, np.uint8))
    nn = RoomNet(num_classes=6, im_side=224, compute_bn_mean_var=False, optimized_inference=True)
    nn.sess = _StubSession()
    xls = classify_im_dir(nn, str(imgs_dir))
    assert xls == str(imgs_dir) + "_classified_results.xls" and os.path.exists(xls)
    out_dir = str(imgs_dir) + "_classified"
    assert sorted(os.listdir(out_dir)) == sorted(CLASS_LABELS)
    placed = {f: lab for lab in CLASS_LABELS for f in os.listdir(os.path.join(out_dir, lab))}
    assert placed["a.png"] == "Backyard" and placed["b.png"] == "Bedroom" and placed["c.png"] == "LivingRoom"
    # overlay=True writes a modified image (text drawn), overlay=False copies the file verbatim
    over = cv2.imread(os.path.join(out_dir, "Bedroom", "b.png"))
    assert over.shape == (300, 400, 3) and (over != 100).any()
    assert nn.sess.batches == [(4, 224, 224, 3)]  # one batched device call instead of four Session.runs
    xls2 = classify_im_dir(nn, str(imgs_dir), overlay=False)
    assert xls2 == xls
    assert open(os.path.join(out_dir, "Bedroom", "b.png"), "rb").read() == open(imgs_dir / "b.png", "rb").read()
    # results table: header + one row per file: name, label, confidence-as-string (reference infer.py:77-78,96-98)
    cells = _read_biff2(xls)
    assert cells[(0, 0)] == "IMAGE_NAME" and cells[(0, 1)] == "PREDICTED_LABEL"
    rows = {cells[(r, 0)]: (cells[(r, 1)], cells[(r, 2)]) for r in range(1, 5)}
    assert rows["c.png"][0] == "LivingRoom" and abs(float(rows["c.png"][1]) - 0.9) < 1e-6


def test_classify_im_dir_unreadable_file_keeps_reference_exception_type(tmp_path):
    imgs_dir = tmp_path / "images"
    imgs_dir.mkdir()
    (imgs_dir / "broken.png").write_bytes(b"not an image")
    nn = RoomNet(num_classes=6, im_side=224, compute_bn_mean_var=False, optimized_inference=True)
    nn.sess = _StubSession()
    with pytest.raises(AttributeError):  # the reference dies with AttributeError in center_crop (network.py:138)
        classify_im_dir(nn, str(imgs_dir))


def _read_biff2(path):
    data = open(path, "rb").read()
    pos, cells = 0, {}
    while pos < len(data):
        rec, ln = struct.unpack_from("<HH", data, pos)
        body = data[pos + 4:pos + 4 + ln]
        if rec == 0x0004:
            row, col = struct.unpack_from("<HH", body, 0)
            n = body[7]
            cells[(row, col)] = body[8:8 + n].decode("latin-1")
        pos += 4 + ln
    return cells


def test_workload_module_matches_the_oracle_generators():
    """bench.py / tools use roomnet_b200.workload (no oracle import on the product side); its generators and shape
    tables must be byte-identical to the oracle's, or the golden fixtures would not describe the benchmark inputs."""
    import numpy as np
    from oracle import roomnet_oracle as orc
    from roomnet_b200 import workload as wl
    for seed in range(12):
        for side in (224, 300):
            assert np.array_equal(wl.synthetic_image(seed, side), orc.synthetic_image(seed, side))
    assert np.array_equal(wl.synthetic_dense0(300), orc.synthetic_dense0(300))
    for side in (224, 300, 600):
        assert wl.spatial_trace(side) == orc.spatial_trace(side) and wl.flat_len(side) == orc.flat_len(side)
    # 4,486,392,000 conv FLOPs per 224x224 image
    assert sum(wl.conv_flops(i) for i in range(10)) == 4_486_392_000
    assert wl.conv_bytes(3) == 210 * 210 * 64 + 205 * 205 * 64 + 215 * 215 * 64
