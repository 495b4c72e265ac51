"""JPEG front end (row n1: cv2.imread of reference infer.py:81 for baseline JPEG files).

CPU tests pin the host half (marker parser + Huffman decoder, roomnet_b200/csrc/jpeg_host.cpp) together with the
numpy restatement of the device half (oracle/jpeg_decode.py) against the real library, cv2.imdecode, bit for bit.
GPU tests compare the CUDA kernels with cv2.imdecode directly and the file call with the decoded-image call.
"""
import os
import struct
import sys

import cv2
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import jpeg_decode as jd  # noqa: E402
from roomnet_b200 import _capi  # noqa: E402

SF = {"444": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444, "422": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_422,
      "420": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_420, "411": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_411,
      "440": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_440}


def photo(h, w, seed=0):
    """Smooth gradients + texture + a saturated patch: exercises every coefficient range and the clamps."""
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:h, 0:w]
    img = np.stack([127 + 100 * np.sin(x / 17.0 + y / 29.0), 127 + 100 * np.cos(x / 11.0 - y / 23.0),
                    (x * 3 + y * 5) % 256], 2).astype(np.float64)
    img += rng.normal(0, 12, img.shape)
    img[h // 3:h // 2, w // 4:w // 2] = [250, 10, 30]
    return np.clip(img, 0, 255).astype(np.uint8)


def encode(img, sf="420", q=90, extra=()):
    ok, enc = cv2.imencode(".jpg", img, [cv2.IMWRITE_JPEG_QUALITY, q, cv2.IMWRITE_JPEG_SAMPLING_FACTOR, SF[sf], *extra])
    assert ok
    return enc.tobytes()


def with_exif_orientation(data, o, big_endian=False):
    """Inserts an APP1 EXIF segment carrying only the orientation tag right after SOI."""
    bo = ">" if big_endian else "<"
    tiff = (b"MM" if big_endian else b"II") + struct.pack(bo + "HI", 42, 8)
    tiff += struct.pack(bo + "H", 1) + struct.pack(bo + "HHIHH", 0x0112, 3, 1, o, 0) + struct.pack(bo + "I", 0)
    seg = b"Exif\0\0" + tiff
    return data[:2] + b"\xff\xe1" + struct.pack(">H", len(seg) + 2) + seg + data[2:]


def with_16bit_quant_tables(data):
    """Rewrites every 8-bit DQT table as a 16-bit one (Pq = 1) with the same values."""
    out = bytearray(data[:2])
    pos = 2
    while True:
        assert data[pos] == 0xFF
        m = data[pos + 1]
        ln = struct.unpack(">H", data[pos + 2:pos + 4])[0]
        seg = data[pos + 4:pos + 2 + ln]
        if m == 0xDB:
            new = bytearray()
            p = 0
            while p < len(seg):
                assert seg[p] >> 4 == 0
                new.append(0x10 | (seg[p] & 15))
                for v in seg[p + 1:p + 65]:
                    new += struct.pack(">H", v)
                p += 65
            out += b"\xff\xdb" + struct.pack(">H", len(new) + 2) + new
        else:
            out += data[pos:pos + 2 + ln]
        pos += 2 + ln
        if m == 0xDA:
            break
    return bytes(out) + data[pos:]


def cpu_decode(data):
    st, coefs = _capi.jpeg_coefficients(data)
    assert st == _capi.JPEG_OK, st
    return jd.decode_from_coefficients(data, coefs)


CASES = [(h, w, sf, q, extra)
         for (h, w) in [(64, 64), (97, 131), (240, 321), (33, 47), (16, 16), (17, 500), (481, 19)]
         for sf in ("444", "422", "420")
         for (q, extra) in [(30, ()), (75, (cv2.IMWRITE_JPEG_OPTIMIZE, 1)), (95, (cv2.IMWRITE_JPEG_RST_INTERVAL, 3)),
                            (100, ())]]


def test_host_decoder_plus_restatement_match_cv2_bit_for_bit():
    for h, w, sf, q, extra in CASES:
        data = encode(photo(h, w, seed=h + w), sf, q, extra)
        ref = cv2.imdecode(np.frombuffer(data, np.uint8), cv2.IMREAD_COLOR)
        got = cpu_decode(data)
        assert got.shape == ref.shape and np.array_equal(got, ref), (h, w, sf, q, extra)


def test_grey_and_16bit_tables_and_all_orientations():
    grey = cv2.cvtColor(photo(100, 150), cv2.COLOR_BGR2GRAY)
    ok, enc = cv2.imencode(".jpg", grey, [cv2.IMWRITE_JPEG_QUALITY, 80])
    assert np.array_equal(cpu_decode(enc.tobytes()), cv2.imdecode(enc, cv2.IMREAD_COLOR))
    data = with_16bit_quant_tables(encode(photo(72, 90), "420", 60))
    assert np.array_equal(cpu_decode(data), cv2.imdecode(np.frombuffer(data, np.uint8), cv2.IMREAD_COLOR))
    base = encode(photo(70, 110), "420", 85)
    plain = cv2.imdecode(np.frombuffer(base, np.uint8), cv2.IMREAD_COLOR)
    for o in range(1, 9):
        data = with_exif_orientation(base, o, big_endian=bool(o & 1))
        ref = cv2.imdecode(np.frombuffer(data, np.uint8), cv2.IMREAD_COLOR)
        if o >= 5:
            assert ref.shape == (110, 70, 3)  # cv2 does apply the tag: this test means something
        if o != 1:
            assert ref.shape != plain.shape or not np.array_equal(ref, plain)
        info = _capi.jpeg_info(data)
        assert info[0] == _capi.JPEG_OK and info[6] == o and (info[2], info[1]) == ref.shape[:2]
        assert np.array_equal(cpu_decode(data), ref), o


def test_scan_preparation_for_the_device_huffman_decoder():
    """The host's share of the device Huffman path: byte stuffing and restart markers removed, restart segments on
    128-byte subsequence boundaries - against a numpy restatement of the same rules (T.81 B.1.1.5, F.1.2.3)."""
    for h, w, sf, q, extra in CASES[::2] + [(300, 400, "420", 97, (cv2.IMWRITE_JPEG_RST_INTERVAL, 1))]:
        data = encode(photo(h, w, seed=h * w), sf, q, extra)
        st, stream, sub_seg, (n_seg, bpm, total_blocks) = _capi.jpeg_prepare_scan(data)
        assert st == _capi.JPEG_OK
        a = np.frombuffer(data, np.uint8)
        sos = data.index(b"\xff\xda")
        start = sos + 2 + int.from_bytes(data[sos + 2:sos + 4], "big")
        scan = a[start:len(a) - 2]  # up to EOI
        ff = np.flatnonzero(scan[:-1] == 0xFF)
        nxt = scan[ff + 1]
        rst = ff[(nxt >= 0xD0) & (nxt <= 0xD7)]
        keep = np.ones(scan.size, bool)
        keep[ff[nxt == 0] + 1] = False  # the stuffed zero after a data 0xFF
        keep[rst] = False
        keep[rst + 1] = False
        bounds = [0, *rst.tolist(), scan.size]
        want, seg_ids = [], []
        for s_i in range(len(bounds) - 1):
            part = scan[bounds[s_i]:bounds[s_i + 1]][keep[bounds[s_i]:bounds[s_i + 1]]]
            pad = (-part.size) % 128
            want.append(np.concatenate([part, np.zeros(pad, np.uint8)]))
            seg_ids += [s_i] * ((part.size + pad) // 128)
        want = np.concatenate(want)
        assert n_seg == len(bounds) - 1
        assert stream.size == want.size and np.array_equal(stream, want), (h, w, sf, q)
        assert np.array_equal(sub_seg, np.array(seg_ids, np.int32))
        hs, vs = {"444": (1, 1), "422": (2, 1), "420": (2, 2)}[sf]
        assert bpm == hs * vs + 2 and total_blocks == bpm * (-(-w // (8 * hs))) * (-(-h // (8 * vs)))
    # a truncated file is left to the host decoder
    assert _capi.jpeg_prepare_scan(data[:len(data) // 2])[0] != _capi.JPEG_OK


def test_files_the_device_path_does_not_take_are_reported_not_mangled():
    img = photo(64, 64)
    ok, prog = cv2.imencode(".jpg", img, [cv2.IMWRITE_JPEG_PROGRESSIVE, 1])
    assert _capi.jpeg_info(prog.tobytes())[0] == _capi.JPEG_UNSUPPORTED
    for sf in ("411", "440"):
        assert _capi.jpeg_info(encode(img, sf))[0] == _capi.JPEG_UNSUPPORTED
    assert _capi.jpeg_info(cv2.imencode(".png", img)[1].tobytes())[0] == _capi.JPEG_CORRUPT  # not a JPEG at all
    assert _capi.jpeg_info(encode(photo(8, 40)))[0] == _capi.JPEG_UNSUPPORTED  # tiny: the decoder's special cases
    data = encode(photo(120, 160), "420", 90)
    for cut in (len(data) // 2, len(data) - 40, 300):
        st, coefs = _capi.jpeg_coefficients(data[:cut])
        assert st in (_capi.JPEG_CORRUPT, _capi.JPEG_UNSUPPORTED) and coefs is None
    rng = np.random.default_rng(3)
    for _ in range(200):  # bit flips in the entropy-coded segment and in the headers: a status, never a crash
        bad = bytearray(data)
        for _ in range(int(rng.integers(1, 6))):
            bad[int(rng.integers(2, len(bad)))] = int(rng.integers(0, 256))
        st, _ = _capi.jpeg_coefficients(bytes(bad))
        assert st in (_capi.JPEG_OK, _capi.JPEG_UNSUPPORTED, _capi.JPEG_CORRUPT)
    assert _capi.jpeg_info(b"\xff\xd8")[0] == _capi.JPEG_CORRUPT
    assert _capi.jpeg_info(b"")[0] == _capi.JPEG_CORRUPT


# ---------------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module", params=["device_huffman", "host_huffman"])
def handle(request):
    """Both halves of the decoder on the device (default), or Huffman decoding on host threads."""
    from roomnet_b200.workload import default_checkpoint_prefix
    h = _capi.Handle(precision="fp16", max_batch=64, jpeg_host_huffman=request.param == "host_huffman")
    h.load_tf_checkpoint(default_checkpoint_prefix())
    h.huffman_on_device = request.param == "device_huffman"
    yield h
    h.close()


@pytest.mark.gpu
def test_device_decode_is_bit_identical_to_cv2(handle):
    files = [encode(photo(h, w, seed=h + w), sf, q, extra) for h, w, sf, q, extra in CASES[::3]]
    files.append(encode(photo(1213, 1777, seed=9), "420", 92))  # a photograph-sized file, odd in both directions
    files.append(encode(photo(900, 1400, seed=10), "422", 85, (cv2.IMWRITE_JPEG_RST_INTERVAL, 7)))
    files.append(cv2.imencode(".jpg", cv2.cvtColor(photo(300, 500), cv2.COLOR_BGR2GRAY))[1].tobytes())
    files.append(with_16bit_quant_tables(encode(photo(72, 90), "420", 60)))
    base = encode(photo(170, 250), "420", 85)
    files += [with_exif_orientation(base, o) for o in range(1, 9)]
    before = handle.jpeg_counters()
    for k, data in enumerate(files):
        ref = cv2.imdecode(np.frombuffer(data, np.uint8), cv2.IMREAD_COLOR)
        got, st = handle.decode_jpeg(data)
        assert st == _capi.JPEG_OK, k
        assert got.shape == ref.shape and np.array_equal(got, ref), k
    dev, host = (a - b for a, b in zip(handle.jpeg_counters(), before))
    # every one of these files is a single interleaved scan: the Huffman decoder that ran is the one asked for
    assert (dev, host) == ((len(files), 0) if handle.huffman_on_device else (0, len(files)))


@pytest.mark.gpu
def test_file_call_equals_decoded_image_call_and_reports_the_rest(handle):
    rng = np.random.default_rng(11)
    files, kinds = [], []
    for k in range(23):
        h, w = int(rng.integers(120, 700)), int(rng.integers(120, 700))
        img = photo(h, w, seed=k)
        kind = k % 6
        if kind == 4:
            data = cv2.imencode(".png", img)[1].tobytes()
        elif kind == 5:
            data = cv2.imencode(".jpg", img, [cv2.IMWRITE_JPEG_PROGRESSIVE, 1])[1].tobytes()
        else:
            data = encode(img, ("444", "422", "420", "420")[kind], int(rng.integers(50, 98)))
            if k % 5 == 0:
                data = with_exif_orientation(data, int(rng.integers(2, 9)))
        files.append(data)
        kinds.append(kind)
    files.append(files[0][:len(files[0]) // 2])  # truncated
    kinds.append(6)
    before = handle.jpeg_counters()
    top1, probs, logits, status = handle.infer_jpeg(files, want_logits=True)
    ok = [i for i, k in enumerate(kinds) if k < 4]
    dev, host = (a - b for a, b in zip(handle.jpeg_counters(), before))
    assert (dev, host) == ((len(ok), 0) if handle.huffman_on_device else (0, len(ok)))
    assert all(status[i] == _capi.JPEG_OK for i in ok)
    assert all(status[i] == _capi.JPEG_UNSUPPORTED for i, k in enumerate(kinds) if k == 5)
    assert all(status[i] != _capi.JPEG_OK for i, k in enumerate(kinds) if k in (4, 6))
    assert all(top1[i] == -1 for i in range(len(files)) if i not in ok)  # untouched
    ims = [cv2.imdecode(np.frombuffer(files[i], np.uint8), cv2.IMREAD_COLOR) for i in ok]
    t2, p2, l2 = handle.infer_images_u8_bgr(ims, want_logits=True)
    assert np.array_equal(logits[ok], l2) and np.array_equal(probs[ok], p2) and np.array_equal(top1[ok], t2)
    # a different thread count and list order change nothing
    t3, p3, l3, s3 = handle.infer_jpeg(files[::-1], threads=1, want_logits=True)
    assert np.array_equal(l3[::-1][ok], l2) and np.array_equal(s3[::-1], status)


@pytest.mark.gpu
def test_damaged_streams_never_break_the_device_decoder():
    """Bytes flipped inside the entropy-coded data (headers intact): the device Huffman decoder must survive whatever
    the stream then says - a status per file, no CUDA error - and where both decoders accept a file they agree."""
    from roomnet_b200.workload import default_checkpoint_prefix
    rng = np.random.default_rng(21)
    good = [encode(photo(160 + 8 * k, 240 - 8 * k, seed=k), ("444", "422", "420")[k % 3], 80,
                   (cv2.IMWRITE_JPEG_RST_INTERVAL, 5) if k % 4 == 0 else ()) for k in range(12)]
    files = []
    for k in range(60):
        d = bytearray(good[k % len(good)])
        start = d.index(b"\xff\xda") + 14
        for _ in range(int(rng.integers(1, 5))):
            mode = int(rng.integers(0, 3))
            pos = int(rng.integers(start, len(d) - 2))
            if mode == 0:
                d[pos] = int(rng.integers(0, 256))
            elif mode == 1:
                del d[pos:pos + int(rng.integers(1, 40))]
            else:
                d[pos:pos] = bytes(rng.integers(0, 256, int(rng.integers(1, 20)), dtype=np.uint8))
        files.append(bytes(d))
    files += good  # undamaged files in the same call still come out right
    outs = {}
    for mode in ("device", "host"):
        h = _capi.Handle(precision="fp16", max_batch=32, jpeg_host_huffman=mode == "host")
        h.load_tf_checkpoint(default_checkpoint_prefix())
        outs[mode] = h.infer_jpeg(files, want_logits=True)
        again = h.infer_jpeg(good, want_logits=True)  # and the handle is still healthy afterwards
        assert (again[3] == 0).all()
        outs[mode + "_good"] = again
        h.close()
    sd, sh = outs["device"][3], outs["host"][3]
    assert set(sd.tolist()) <= {0, 1, 2} and set(sh.tolist()) <= {0, 1, 2}
    assert (sd[-len(good):] == 0).all() and (sh[-len(good):] == 0).all()
    both = (sd == 0) & (sh == 0)
    assert both.sum() >= len(good)
    assert np.array_equal(outs["device"][2][both], outs["host"][2][both])
    assert np.array_equal(outs["device_good"][2], outs["host_good"][2])
    assert np.array_equal(outs["device"][2][-len(good):], outs["device_good"][2])


@pytest.mark.gpu
def test_infer_files_falls_back_to_cv2_for_what_the_device_does_not_decode():
    from roomnet_b200.network import RoomNet
    from roomnet_b200.workload import default_checkpoint_prefix
    nn = RoomNet(num_classes=6, im_side=224, compute_bn_mean_var=False, optimized_inference=True)
    nn.load(default_checkpoint_prefix())
    imgs = [photo(200 + 37 * k, 380 - 21 * k, seed=k) for k in range(7)]
    blobs = [encode(imgs[0], "420", 90), cv2.imencode(".png", imgs[1])[1].tobytes(),
             cv2.imencode(".jpg", imgs[2], [cv2.IMWRITE_JPEG_PROGRESSIVE, 1])[1].tobytes(), encode(imgs[3], "444", 70),
             encode(imgs[4], "411", 80), with_exif_orientation(encode(imgs[5], "422", 88), 6), encode(imgs[6], "420", 55)]
    top1, probs = nn.infer_files(blobs)
    for i, b in enumerate(blobs):
        t, p = nn.infer_optimized(cv2.imdecode(np.frombuffer(b, np.uint8), cv2.IMREAD_COLOR))
        assert t[0] == top1[i] and np.array_equal(p[0], probs[i]), i
    with pytest.raises(AttributeError):
        nn.infer_files([blobs[0], b"not an image"])
    nn.close()
