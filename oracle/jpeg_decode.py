"""ORACLE / TEST INFRASTRUCTURE — not part of the product path.

CPU restatement of the second half of the JPEG decoder that stands behind the reference's ``cv2.imread``
(/root/reference/infer.py:81): dequantisation, inverse DCT, chroma upsampling, colour conversion and EXIF orientation,
in numpy integer arithmetic.  The algorithm lives in a third-party dependency that is not vendored in the reference:
OpenCV's bundled **libjpeg-turbo (3.1.2 in the opencv-python 4.13 wheel of this image)** with its default settings —
``JDCT_ISLOW`` (jidctint.c: Loeffler-Ligtenberg-Moschytz, 13-bit constants, 2 pass bits), "fancy" triangle-filter
upsampling (jdsample.c h2v1/h2v2_fancy_upsample, context rows replicated at the image border, jdmainct.c) and the
16-bit fixed-point YCbCr->RGB tables (jdcolor.c); OpenCV then applies the EXIF orientation (imgcodecs/loadsave.cpp
ExifTransform).  This restatement follows those published algorithms; it is PINNED against the real library:
tests/test_jpeg.py compares it bit for bit with ``cv2.imdecode`` on every sampling mode, odd sizes, restart
intervals, optimised Huffman tables, 16-bit quantisation tables and all eight orientations.

The first half (marker parsing + Huffman decoding) is host code of the product (roomnet_b200/csrc/jpeg_host.cpp,
ITU-T T.81); the tests feed ITS coefficients through THIS restatement and compare with cv2, which pins both halves on
the CPU; the GPU tests compare the device kernels with cv2 directly.
"""
from __future__ import annotations

import numpy as np

ZIGZAG = np.array([0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13,
                   6, 7, 14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38,
                   31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63])


def parse_header(data: bytes) -> dict:
    """Frame geometry, quantisation tables (natural order) and EXIF orientation — Annex B of T.81."""
    assert data[:2] == b"\xff\xd8", "not a JPEG"
    pos = 2
    quant = {}
    out = {"orientation": 1}
    while pos < len(data):
        while data[pos] != 0xFF:
            pos += 1
        while data[pos] == 0xFF:
            pos += 1
        m = data[pos]
        pos += 1
        if m in (0x01,) or 0xD0 <= m <= 0xD7:
            continue
        ln = int.from_bytes(data[pos:pos + 2], "big")
        seg = data[pos + 2:pos + ln]
        pos += ln
        if m == 0xDB:
            p = 0
            while p < len(seg):
                pq, tq = seg[p] >> 4, seg[p] & 15
                if pq:
                    vals = np.frombuffer(seg[p + 1:p + 129], dtype=">u2").astype(np.int64)
                    p += 129
                else:
                    vals = np.frombuffer(seg[p + 1:p + 65], dtype=np.uint8).astype(np.int64)
                    p += 65
                t = np.zeros(64, np.int64)
                t[ZIGZAG] = vals
                quant[tq] = t.reshape(8, 8)
        elif m in (0xC0, 0xC1):
            out["height"] = int.from_bytes(seg[1:3], "big")
            out["width"] = int.from_bytes(seg[3:5], "big")
            out["comps"] = [dict(id=seg[6 + 3 * c], h=seg[7 + 3 * c] >> 4, v=seg[7 + 3 * c] & 15, tq=seg[8 + 3 * c])
                            for c in range(seg[5])]
        elif m == 0xE1 and seg[:6] == b"Exif\0\0" and "exif_seen" not in out:
            out["exif_seen"] = True
            t = seg[6:]
            bo = "little" if t[:2] == b"II" else "big"
            ifd = int.from_bytes(t[4:8], bo)
            for i in range(int.from_bytes(t[ifd:ifd + 2], bo)):
                e = ifd + 2 + 12 * i
                if int.from_bytes(t[e:e + 2], bo) == 0x0112:
                    o = int.from_bytes(t[e + 8:e + 10], bo)
                    out["orientation"] = o if 1 <= o <= 8 else 1
        elif m == 0xDA:
            break
    out["quant"] = quant
    if len(out["comps"]) == 1:
        out["comps"][0]["h"] = out["comps"][0]["v"] = 1
    hmax = max(c["h"] for c in out["comps"])
    vmax = max(c["v"] for c in out["comps"])
    out["hmax"], out["vmax"] = hmax, vmax
    for c in out["comps"]:
        c["dw"] = -(-out["width"] * c["h"] // hmax)
        c["dh"] = -(-out["height"] * c["v"] // vmax)
        c["wblocks"] = -(-c["dw"] // 8)
        c["hblocks"] = -(-c["dh"] // 8)
    return out


def _idct8(i0, i1, i2, i3, i4, i5, i6, i7):
    """One 8-point pass of jpeg_idct_islow before the descale (int64 arrays)."""
    z2, z3 = i2, i6
    z1 = (z2 + z3) * 4433
    tmp2 = z1 - z3 * 15137
    tmp3 = z1 + z2 * 6270
    tmp0 = (i0 + i4) << 13
    tmp1 = (i0 - i4) << 13
    tmp10, tmp13, tmp11, tmp12 = tmp0 + tmp3, tmp0 - tmp3, tmp1 + tmp2, tmp1 - tmp2
    tmp0, tmp1, tmp2, tmp3 = i7, i5, i3, i1
    z1, z2, z3, z4 = tmp0 + tmp3, tmp1 + tmp2, tmp0 + tmp2, tmp1 + tmp3
    z5 = (z3 + z4) * 9633
    tmp0, tmp1, tmp2, tmp3 = tmp0 * 2446, tmp1 * 16819, tmp2 * 25172, tmp3 * 12299
    z1, z2, z3, z4 = z1 * -7373, z2 * -20995, z3 * -16069 + z5, z4 * -3196 + z5
    tmp0, tmp1, tmp2, tmp3 = tmp0 + z1 + z3, tmp1 + z2 + z4, tmp2 + z2 + z3, tmp3 + z1 + z4
    return (tmp10 + tmp3, tmp11 + tmp2, tmp12 + tmp1, tmp13 + tmp0, tmp13 - tmp0, tmp12 - tmp1, tmp11 - tmp2,
            tmp10 - tmp3)


def idct_islow(coefs: np.ndarray, quant: np.ndarray) -> np.ndarray:
    """coefs [hb, wb, 8, 8] quantised coefficients (natural order) -> sample plane [hb*8, wb*8] uint8."""
    x = coefs.astype(np.int64) * quant[None, None]
    cols = _idct8(*[x[:, :, r, :] for r in range(8)])  # pass 1: down the columns
    ws = np.stack([(c + (1 << 10)) >> 11 for c in cols], axis=2)  # [hb, wb, r, c]
    rows = _idct8(*[ws[:, :, :, k] for k in range(8)])  # pass 2: along the rows
    v = np.stack([(r + (1 << 17)) >> 18 for r in rows], axis=3)  # [hb, wb, r, c]
    idx = v & 1023  # the decoder's range-limit table (centre +128, clamp, wrap outside +-512)
    s = np.where(idx < 128, 128 + idx, np.where(idx < 512, 255, np.where(idx < 896, 0, idx - 896)))
    hb, wb = s.shape[:2]
    return s.transpose(0, 2, 1, 3).reshape(hb * 8, wb * 8).astype(np.uint8)


def _fancy_h(rows_near: np.ndarray, rows_far, dw: int, out_w: int) -> np.ndarray:
    """Horizontal triangle filter of jdsample.c; rows_far is None for h2v1, else the further row (h2v2)."""
    a = rows_near[:, :dw].astype(np.int64)
    if rows_far is None:
        cur = a
        prev = np.concatenate([a[:, :1], a[:, :-1]], axis=1)
        nxt = np.concatenate([a[:, 1:], a[:, -1:]], axis=1)
        even = (3 * cur + prev + 1) >> 2
        odd = (3 * cur + nxt + 2) >> 2
        even[:, 0] = cur[:, 0]
        odd[:, -1] = cur[:, -1]
    else:
        cur = 3 * a + rows_far[:, :dw].astype(np.int64)
        prev = np.concatenate([cur[:, :1], cur[:, :-1]], axis=1)
        nxt = np.concatenate([cur[:, 1:], cur[:, -1:]], axis=1)
        even = (3 * cur + prev + 8) >> 4
        odd = (3 * cur + nxt + 7) >> 4
        even[:, 0] = (4 * cur[:, 0] + 8) >> 4
        odd[:, -1] = (4 * cur[:, -1] + 7) >> 4
    out = np.empty((a.shape[0], 2 * dw), np.int64)
    out[:, 0::2] = even
    out[:, 1::2] = odd
    return out[:, :out_w]


def upsample(plane: np.ndarray, hs: int, vs: int, dw: int, dh: int, W: int, H: int) -> np.ndarray:
    if hs == 1 and vs == 1:
        return plane[:H, :W].astype(np.int64)
    if vs == 1:
        return _fancy_h(plane[:H], None, dw, W)
    y = np.arange(H)
    r0 = y >> 1
    r1 = np.where(y & 1, np.minimum(r0 + 1, dh - 1), np.maximum(r0 - 1, 0))
    return _fancy_h(plane[r0], plane[r1], dw, W)


def orient(img: np.ndarray, o: int) -> np.ndarray:
    """OpenCV's ExifTransform (modules/imgcodecs/src/loadsave.cpp)."""
    if o == 2:
        return img[:, ::-1]
    if o == 3:
        return img[::-1, ::-1]
    if o == 4:
        return img[::-1]
    if o == 5:
        return img.transpose(1, 0, 2)
    if o == 6:
        return img.transpose(1, 0, 2)[:, ::-1]
    if o == 7:
        return img.transpose(1, 0, 2)[::-1, ::-1]
    if o == 8:
        return img.transpose(1, 0, 2)[::-1]
    return img


def decode_from_coefficients(data: bytes, coefs: np.ndarray) -> np.ndarray:
    """BGR uint8 image from the file's header and its entropy-decoded coefficients (components back to back)."""
    hd = parse_header(data)
    W, H = hd["width"], hd["height"]
    planes = []
    off = 0
    for c in hd["comps"]:
        n = c["wblocks"] * c["hblocks"] * 64
        blk = np.asarray(coefs[off:off + n]).reshape(c["hblocks"], c["wblocks"], 8, 8)
        off += n
        planes.append(idct_islow(blk, hd["quant"][c["tq"]]))
    Y = planes[0][:H, :W].astype(np.int64)
    if len(planes) == 1:
        img = np.stack([Y, Y, Y], axis=2)
    else:
        c1 = hd["comps"][1]
        cb = upsample(planes[1], hd["hmax"], hd["vmax"], c1["dw"], c1["dh"], W, H) - 128
        cr = upsample(planes[2], hd["hmax"], hd["vmax"], c1["dw"], c1["dh"], W, H) - 128
        r = Y + ((91881 * cr + 32768) >> 16)
        b = Y + ((116130 * cb + 32768) >> 16)
        g = Y + ((-22554 * cb + 32768 - 46802 * cr) >> 16)
        img = np.stack([b, g, r], axis=2)
    img = np.clip(img, 0, 255).astype(np.uint8)
    return np.ascontiguousarray(orient(img, hd["orientation"]))
