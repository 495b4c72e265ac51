"""ORACLE / TEST INFRASTRUCTURE — not part of the product path.

Pure-Python reader for TensorFlow "V2 bundle" checkpoints
(``<prefix>.index`` + ``<prefix>.data-00000-of-00001``), the format the
reference writes with ``tf.train.Saver`` (reference network.py:46-47, :94-97)
and restores with ``Saver.restore`` (reference network.py:122).

TensorFlow 1.13.1 itself is not installable here, so the on-disk format is
restated from its published layout:

* ``.index`` is a LevelDB-style SSTable (48-byte footer with the metaindex
  and index block handles as varint64 pairs and the magic
  0xdb4775248b80fb57; blocks are prefix-compressed key/value entries followed
  by a restart array, a 1-byte compression tag and a 4-byte masked CRC32C).
* key ``""`` holds ``BundleHeaderProto``; every other key is a tensor name whose
  value is a ``BundleEntryProto`` {1: dtype, 2: shape, 3: shard_id, 4: offset,
  5: size, 6: masked crc32c (fixed32)}.
* ``.data-*`` is the raw little-endian tensor bytes at ``offset``.

The product's C++ reader (roomnet_b200/csrc/tf_bundle.cpp) is an independent
implementation; tests compare the two.
"""
from __future__ import annotations

import os
import struct
from dataclasses import dataclass

import numpy as np

TABLE_MAGIC = 0xDB4775248B80FB57
DT_FLOAT = 1
_DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 9: np.int64}


# ----------------------------------------------------------------- CRC32C --
def _make_crc_table():
    poly = 0x82F63B78
    tab = []
    for i in range(256):
        c = i
        for _ in range(8):
            c = (c >> 1) ^ poly if c & 1 else c >> 1
        tab.append(c)
    return np.array(tab, dtype=np.uint32)


_CRC_TABLE = _make_crc_table()


def crc32c(data: bytes) -> int:
    """Castagnoli CRC (reflected poly 0x82F63B78), bytewise table."""
    tab = _CRC_TABLE
    c = 0xFFFFFFFF
    for b in data:
        c = int(tab[(c ^ b) & 0xFF]) ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def mask_crc(c: int) -> int:
    """LevelDB/TF CRC mask: rotate right 15 and add a constant."""
    return ((((c >> 15) | (c << 17)) & 0xFFFFFFFF) + 0xA282EAD8) & 0xFFFFFFFF


# ---------------------------------------------------------------- varints --
def _varint(buf: bytes, pos: int):
    shift = 0
    out = 0
    while True:
        b = buf[pos]
        pos += 1
        out |= (b & 0x7F) << shift
        if not b & 0x80:
            return out, pos
        shift += 7


def _parse_proto(buf: bytes):
    """Minimal protobuf wire parser → list of (field, wiretype, value)."""
    pos = 0
    out = []
    while pos < len(buf):
        tag, pos = _varint(buf, pos)
        field, wt = tag >> 3, tag & 7
        if wt == 0:
            v, pos = _varint(buf, pos)
        elif wt == 1:
            v = struct.unpack_from("<Q", buf, pos)[0]
            pos += 8
        elif wt == 2:
            ln, pos = _varint(buf, pos)
            v = buf[pos:pos + ln]
            pos += ln
        elif wt == 5:
            v = struct.unpack_from("<I", buf, pos)[0]
            pos += 4
        else:
            raise ValueError("unsupported protobuf wire type %d" % wt)
        out.append((field, wt, v))
    return out


# ---------------------------------------------------------------- SSTable --
def _read_block(data: bytes, offset: int, size: int, verify: bool = True) -> bytes:
    body = data[offset:offset + size]
    ctype = data[offset + size]
    if ctype != 0:
        raise ValueError("compressed SSTable blocks are not supported (type %d)" % ctype)
    if verify:
        stored = struct.unpack_from("<I", data, offset + size + 1)[0]
        actual = mask_crc(crc32c(data[offset:offset + size + 1]))
        if stored != actual:
            raise ValueError("SSTable block CRC mismatch")
    return body


def _block_entries(block: bytes):
    n_restarts = struct.unpack_from("<I", block, len(block) - 4)[0]
    end = len(block) - 4 - 4 * n_restarts
    pos = 0
    key = b""
    while pos < end:
        shared, pos = _varint(block, pos)
        non_shared, pos = _varint(block, pos)
        vlen, pos = _varint(block, pos)
        key = key[:shared] + block[pos:pos + non_shared]
        pos += non_shared
        val = block[pos:pos + vlen]
        pos += vlen
        yield key, val


@dataclass
class BundleEntry:
    name: str
    dtype: int
    shape: tuple
    shard_id: int
    offset: int
    size: int
    crc32c: int


def read_index(index_path: str):
    """Returns (header dict, {name: BundleEntry}) in lexicographic key order."""
    data = open(index_path, "rb").read()
    if len(data) < 48:
        raise ValueError("index file too short")
    footer = data[-48:]
    if struct.unpack_from("<Q", footer, 40)[0] != TABLE_MAGIC:
        raise ValueError("bad SSTable magic")
    pos = 0
    _mi_off, pos = _varint(footer, pos)
    _mi_size, pos = _varint(footer, pos)
    idx_off, pos = _varint(footer, pos)
    idx_size, pos = _varint(footer, pos)
    index_block = _read_block(data, idx_off, idx_size)
    header = {}
    entries = {}
    for _, handle in _block_entries(index_block):
        off, p = _varint(handle, 0)
        size, p = _varint(handle, p)
        for key, val in _block_entries(_read_block(data, off, size)):
            fields = _parse_proto(val)
            if key == b"":
                for f, _, v in fields:
                    if f == 1:
                        header["num_shards"] = v
                    elif f == 2:
                        header["endianness"] = v
                header.setdefault("num_shards", 0)
                header.setdefault("endianness", 0)
                continue
            e = dict(dtype=0, shape=(), shard_id=0, offset=0, size=0, crc32c=0)
            for f, _, v in fields:
                if f == 1:
                    e["dtype"] = v
                elif f == 2:
                    dims = []
                    for f2, _, v2 in _parse_proto(v):
                        if f2 == 2:
                            sz = 0
                            for f3, _, v3 in _parse_proto(v2):
                                if f3 == 1:
                                    sz = v3
                            dims.append(sz)
                    e["shape"] = tuple(dims)
                elif f == 3:
                    e["shard_id"] = v
                elif f == 4:
                    e["offset"] = v
                elif f == 5:
                    e["size"] = v
                elif f == 6:
                    e["crc32c"] = v
            name = key.decode("utf-8")
            entries[name] = BundleEntry(name=name, **e)
    return header, entries


def load_checkpoint(prefix: str, verify_crc: bool = True):
    """Reads every tensor of ``<prefix>.index/.data-00000-of-00001``.

    Returns {name: np.ndarray}. Mirrors what ``Saver.restore`` (reference
    network.py:122) makes available to the graph.
    """
    header, entries = read_index(prefix + ".index")
    if header.get("num_shards", 1) != 1:
        raise ValueError("only single-shard bundles are supported")
    if header.get("endianness", 0) != 0:
        raise ValueError("big-endian bundles are not supported")
    data_path = prefix + ".data-00000-of-00001"
    blob = open(data_path, "rb").read()
    out = {}
    for name, e in entries.items():
        if e.dtype not in _DTYPES:
            raise ValueError("unsupported dtype %d for %s" % (e.dtype, name))
        raw = blob[e.offset:e.offset + e.size]
        if len(raw) != e.size:
            raise ValueError("tensor %s runs past the end of the data file" % name)
        if verify_crc and mask_crc(crc32c(raw)) != e.crc32c:
            raise ValueError("tensor %s CRC32C mismatch" % name)
        out[name] = np.frombuffer(raw, dtype=_DTYPES[e.dtype]).reshape(e.shape).copy()
    return out


def default_checkpoint_prefix() -> str:
    """Location of the shipped ``final_model`` weights.

    ``/root/reference`` exists only in the build container; the GPU box uses the
    byte-identical copy under final_model/ at the repository root (the weights are
    data, not source; sha256 pinned in tests/test_bundle.py).
    """
    here = os.path.dirname(os.path.abspath(__file__))
    fixture = os.path.join(here, "..", "final_model", "roomnet")
    if os.path.exists(fixture + ".index"):
        return os.path.normpath(fixture)
    return "/root/reference/final_model/roomnet"
