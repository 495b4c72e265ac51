"""Checkpoint reader known-answers for the oracle reader and the product's C++ reader."""
import hashlib
import os
import shutil

import numpy as np
import pytest

from oracle import tf_bundle


def test_fixture_is_the_shipped_checkpoint(ckpt_prefix):
    sha = lambda p: hashlib.sha256(open(p, "rb").read()).hexdigest()  # noqa: E731
    assert sha(ckpt_prefix + ".data-00000-of-00001") == "deb5beefd23b333e638a1b91c82a20b6bc4c871cb0acf45bee5896a0a41cc5d5"
    assert sha(ckpt_prefix + ".index") == "9a1186c136984a00bfc76f2ba5e1935354390f8f6b9c8469a7ddb7e1537b3a3e"


def test_crc32c_known_answers():
    assert tf_bundle.crc32c(b"123456789") == 0xE3069283  # standard CRC-32C check value
    assert tf_bundle.crc32c(b"") == 0


def test_index_entries(ckpt_prefix):
    header, entries = tf_bundle.read_index(ckpt_prefix + ".index")
    assert header["num_shards"] == 1
    assert len(entries) == 79
    assert sum(e.size for e in entries.values()) == 712248
    assert entries["batch_normalization/beta"].crc32c == 0x0F41B09C
    assert entries["conv2d/kernel"].crc32c == 0x652BBE26
    assert entries["dense_3/bias"].crc32c == 0xB39D390F
    assert entries["dense_3/kernel"].crc32c == 0x7BEF4B98
    expect = {"conv2d/kernel": ((3, 3, 3, 8), 9216), "conv2d_1/kernel": ((3, 3, 8, 32), 10080),
              "conv2d_4/kernel": ((3, 3, 32, 64), 93024), "conv2d_6/kernel": ((3, 3, 64, 128), 314208),
              "conv2d_7/kernel": ((3, 3, 128, 16), 609120), "dense/kernel": ((64, 32), 701280),
              "dense_3/bias": ((6,), 712032), "dense_3/kernel": ((8, 6), 712056)}
    for name, (shape, off) in expect.items():
        assert entries[name].shape == shape and entries[name].offset == off, name
    assert list(entries) == sorted(entries)  # lexicographic key order


def test_tensor_values(weights):
    assert len(weights) == 79 and sum(v.size for v in weights.values()) == 178062
    np.testing.assert_allclose(weights["dense_3/bias"],
                               [0.07939394, 0.10200979, 0.08301026, 0.05211717, 0.14262062, 0.11053859], rtol=1e-6)
    np.testing.assert_allclose(weights["dense_3/kernel"][0],
                               [-0.17933735, 0.29178998, -0.5602683, -0.35227767, 0.31789997, 0.517098], rtol=1e-6)


def test_crc_mismatch_is_detected(ckpt_prefix, tmp_path):
    for ext in (".index", ".data-00000-of-00001"):
        shutil.copy(ckpt_prefix + ext, tmp_path / ("roomnet" + ext))
    data = tmp_path / "roomnet.data-00000-of-00001"
    blob = bytearray(data.read_bytes())
    blob[20000] ^= 0x40
    data.write_bytes(bytes(blob))
    with pytest.raises(ValueError, match="CRC32C"):
        tf_bundle.load_checkpoint(str(tmp_path / "roomnet"))


# ---- the product's C++ reader + folder, through the C ABI on a host-only handle ----------------
def _host_handle(capi, **kw):
    return capi.Handle(devices=(), **kw)


def test_cpp_reader_and_fold_match_oracle_fold(capi, ckpt_prefix, weights):
    from oracle.fold import fold
    h = _host_handle(capi)
    h.load_tf_checkpoint(ckpt_prefix)
    F = fold(weights)
    Frgb = fold(weights, u8_bgr_input=False)
    np.testing.assert_allclose(h.get_folded("conv0_u8bgr/w"), F["convs"][0]["W"].ravel(), rtol=1e-6, atol=1e-12)
    np.testing.assert_allclose(h.get_folded("conv0_u8bgr/b"), F["convs"][0]["b"], rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(h.get_folded("conv0_f32rgb/w"), Frgb["convs"][0]["W"].ravel(), rtol=1e-6, atol=1e-12)
    assert np.all(h.get_folded("conv0_f32rgb/b") == 0)
    # u8 RGB = the BGR fold with the channel axis reversed
    w_rgb = h.get_folded("conv0_u8rgb/w").reshape(3, 3, 3, 8)
    np.testing.assert_allclose(w_rgb[:, :, ::-1, :], F["convs"][0]["W"], rtol=1e-6, atol=1e-12)
    for i in range(1, 10):
        np.testing.assert_allclose(h.get_folded("conv%d/w" % i), F["convs"][i]["W"].ravel(), rtol=1e-6, atol=1e-12)
        np.testing.assert_allclose(h.get_folded("conv%d/b" % i), F["convs"][i]["b"], rtol=1e-6, atol=1e-9)
    assert sorted(F["joins"]) == [3, 5, 9]
    for i, (A, B, Cc, src) in F["joins"].items():
        np.testing.assert_allclose(h.get_folded("join%d/a" % i), A, rtol=1e-6)
        np.testing.assert_allclose(h.get_folded("join%d/b" % i), B, rtol=1e-6)
        np.testing.assert_allclose(h.get_folded("join%d/c" % i), Cc, rtol=1e-6, atol=1e-9)
    for i in range(4):
        np.testing.assert_allclose(h.get_folded("dense%d/w" % i), F["dense"][i]["W"].ravel(), rtol=1e-6, atol=1e-12)
        np.testing.assert_allclose(h.get_folded("dense%d/b" % i), F["dense"][i]["b"], rtol=1e-6, atol=1e-9)
    # layers that consume a residual join carry no folded bias
    assert np.all(h.get_folded("conv4/b") == 0) and np.all(h.get_folded("conv6/b") == 0)
    assert np.all(h.get_folded("dense0/b") == 0)
    h.close()


def test_cpp_load_tensors_equals_load_checkpoint(capi, ckpt_prefix, weights):
    a, b = _host_handle(capi), _host_handle(capi)
    a.load_tf_checkpoint(ckpt_prefix)
    b.load_tensors(weights)
    for name in ("conv5/w", "join9/c", "dense3/b"):
        assert np.array_equal(a.get_folded(name), b.get_folded(name))


def test_cpp_error_codes(capi, ckpt_prefix, tmp_path):
    h = _host_handle(capi)
    with pytest.raises(capi.RoomNetError) as e:
        h.load_tf_checkpoint(str(tmp_path / "missing"))
    assert e.value.code == capi.RN_ERR_IO
    for ext in (".index", ".data-00000-of-00001"):
        shutil.copy(ckpt_prefix + ext, tmp_path / ("roomnet" + ext))
    data = tmp_path / "roomnet.data-00000-of-00001"
    blob = bytearray(data.read_bytes())
    blob[300000] ^= 0x01
    data.write_bytes(bytes(blob))
    with pytest.raises(capi.RoomNetError) as e:
        h.load_tf_checkpoint(str(tmp_path / "roomnet"))
    assert e.value.code == capi.RN_ERR_FORMAT and "CRC32C" in str(e.value)
    idx = tmp_path / "roomnet.index"
    iblob = bytearray(idx.read_bytes())
    iblob[100] ^= 0xFF
    idx.write_bytes(bytes(iblob))
    with pytest.raises(capi.RoomNetError) as e:
        h.load_tf_checkpoint(str(tmp_path / "roomnet"))
    assert e.value.code == capi.RN_ERR_FORMAT
    # the shipped dense/kernel only fits im_side 224
    h300 = _host_handle(capi, im_side=300)
    assert h300.flat_len == 256
    with pytest.raises(capi.RoomNetError) as e:
        h300.load_tf_checkpoint(ckpt_prefix)
    assert e.value.code == capi.RN_ERR_FORMAT and "dense/kernel" in str(e.value)
    h300.set_dense0(np.zeros((256, 32), np.float32))
    h300.load_tf_checkpoint(ckpt_prefix)
    with pytest.raises(capi.RoomNetError):
        h300.set_dense0(np.zeros((64, 32), np.float32))


def _varint(v):
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        out.append(b | (0x80 if v else 0))
        if not v:
            return bytes(out)


def test_cpp_reader_rejects_overflowing_handles(capi, ckpt_prefix, tmp_path):
    """Untrusted 64-bit block handles / entry lengths must not wrap the bounds checks (off + size + 5)."""
    h = _host_handle(capi)
    for ext in (".index", ".data-00000-of-00001"):
        shutil.copy(ckpt_prefix + ext, tmp_path / ("roomnet" + ext))
    idx = tmp_path / "roomnet.index"
    good = idx.read_bytes()
    magic = good[-8:]
    for off, size in ((1 << 40, (1 << 64) - (1 << 40) - 5), (0, (1 << 64) - 1), ((1 << 64) - 1, 0), (len(good), 1)):
        footer = _varint(0) + _varint(0) + _varint(off) + _varint(size)
        footer += b"\0" * (40 - len(footer)) + magic
        idx.write_bytes(good[:-48] + footer)
        with pytest.raises(capi.RoomNetError) as e:
            h.load_tf_checkpoint(str(tmp_path / "roomnet"))
        assert e.value.code == capi.RN_ERR_FORMAT, (off, size)
    idx.write_bytes(good)
    h.load_tf_checkpoint(str(tmp_path / "roomnet"))  # the untouched copy still loads
