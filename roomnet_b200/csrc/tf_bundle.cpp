// TensorFlow V2 bundle reader.  Format (TF 1.13.1, tensor_bundle + leveldb table):
//   .index  : SSTable; footer = 2 block handles (varint64 offset,size) padded to
//             40 bytes + magic 0xdb4775248b80fb57.  Blocks: prefix-compressed
//             entries, restart array, 1-byte compression tag, masked CRC32C.
//             key ""  -> BundleHeaderProto {1:num_shards, 2:endianness}
//             key name-> BundleEntryProto  {1:dtype, 2:shape, 3:shard, 4:offset, 5:size, 6:crc32c}
//   .data-* : raw little-endian tensor bytes.
#include "tf_bundle.h"

#include <cstdio>
#include <cstring>

namespace rn {
namespace {

constexpr uint64_t kTableMagic = 0xdb4775248b80fb57ull;

struct Crc32cTable {
  uint32_t t[8][256];
  Crc32cTable() {
    for (uint32_t i = 0; i < 256; ++i) {
      uint32_t c = i;
      for (int k = 0; k < 8; ++k) c = (c & 1) ? (c >> 1) ^ 0x82F63B78u : c >> 1;
      t[0][i] = c;
    }
    for (uint32_t i = 0; i < 256; ++i)
      for (int s = 1; s < 8; ++s) t[s][i] = (t[s - 1][i] >> 8) ^ t[0][t[s - 1][i] & 0xff];
  }
};

const Crc32cTable& Table() {
  static Crc32cTable tab;
  return tab;
}

bool ReadFile(const std::string& path, std::vector<uint8_t>* out) {
  FILE* f = std::fopen(path.c_str(), "rb");
  if (!f) return false;
  std::fseek(f, 0, SEEK_END);
  long n = std::ftell(f);
  std::fseek(f, 0, SEEK_SET);
  if (n < 0) {
    std::fclose(f);
    return false;
  }
  out->resize(static_cast<size_t>(n));
  size_t got = n ? std::fread(out->data(), 1, static_cast<size_t>(n), f) : 0;
  std::fclose(f);
  return got == static_cast<size_t>(n);
}

struct Cursor {
  const uint8_t* p;
  const uint8_t* end;
  bool ok = true;
  uint64_t Varint() {
    uint64_t v = 0;
    for (int shift = 0; shift < 64; shift += 7) {
      if (p >= end) {
        ok = false;
        return 0;
      }
      uint8_t b = *p++;
      v |= static_cast<uint64_t>(b & 0x7f) << shift;
      if (!(b & 0x80)) return v;
    }
    ok = false;
    return 0;
  }
};

struct Slice {
  const uint8_t* p = nullptr;
  size_t n = 0;
};

// Returns the block body after verifying compression tag and CRC.
bool GetBlock(const std::vector<uint8_t>& file, uint64_t off, uint64_t size, Slice* out, std::string* err) {
  // overflow-safe: the handle is untrusted (two 64-bit varints), so never add before comparing
  if (file.size() < 5 || size > file.size() - 5 || off > file.size() - 5 - size) {
    *err = "SSTable block handle runs past the end of the index file";
    return false;
  }
  if (file[off + size] != 0) {
    *err = "compressed SSTable blocks are not supported";
    return false;
  }
  uint32_t stored;
  std::memcpy(&stored, &file[off + size + 1], 4);
  if (MaskCrc(Crc32c(&file[off], size + 1)) != stored) {
    *err = "SSTable block CRC mismatch";
    return false;
  }
  out->p = &file[off];
  out->n = size;
  return true;
}

template <typename Fn>
bool ForEachEntry(Slice block, Fn fn, std::string* err) {
  if (block.n < 4) {
    *err = "SSTable block too small";
    return false;
  }
  uint32_t n_restarts;
  std::memcpy(&n_restarts, block.p + block.n - 4, 4);
  if (4ull + 4ull * n_restarts > block.n) {
    *err = "bad restart array";
    return false;
  }
  Cursor c{block.p, block.p + block.n - 4 - 4ull * n_restarts};
  std::string key;
  while (c.p < c.end) {
    uint64_t shared = c.Varint(), non_shared = c.Varint(), vlen = c.Varint();
    const uint64_t rem = static_cast<uint64_t>(c.end - c.p);
    if (!c.ok || shared > key.size() || non_shared > rem || vlen > rem - non_shared) {
      *err = "corrupt SSTable entry";
      return false;
    }
    key.resize(shared);
    key.append(reinterpret_cast<const char*>(c.p), non_shared);
    c.p += non_shared;
    Slice val{c.p, static_cast<size_t>(vlen)};
    c.p += vlen;
    if (!fn(key, val)) return false;
  }
  return true;
}

struct Entry {
  int dtype = 0;
  std::vector<int64_t> shape;
  int64_t shard = 0, offset = 0, size = 0;
  uint32_t crc = 0;
};

// Minimal protobuf walk: calls fn(field, wiretype, varint_or_fixed, bytes).
template <typename Fn>
bool WalkProto(Slice s, Fn fn) {
  Cursor c{s.p, s.p + s.n};
  while (c.p < c.end) {
    uint64_t tag = c.Varint();
    if (!c.ok) return false;
    int field = static_cast<int>(tag >> 3), wt = static_cast<int>(tag & 7);
    uint64_t v = 0;
    Slice b;
    if (wt == 0) {
      v = c.Varint();
    } else if (wt == 1) {
      if (c.end - c.p < 8) return false;
      std::memcpy(&v, c.p, 8);
      c.p += 8;
    } else if (wt == 2) {
      uint64_t len = c.Varint();
      if (!c.ok || static_cast<uint64_t>(c.end - c.p) < len) return false;
      b = Slice{c.p, static_cast<size_t>(len)};
      c.p += len;
    } else if (wt == 5) {
      if (c.end - c.p < 4) return false;
      uint32_t v32;
      std::memcpy(&v32, c.p, 4);
      v = v32;
      c.p += 4;
    } else {
      return false;
    }
    if (!c.ok) return false;
    fn(field, wt, v, b);
  }
  return true;
}

bool ParseEntry(Slice val, Entry* e) {
  return WalkProto(val, [&](int f, int, uint64_t v, Slice b) {
    switch (f) {
      case 1: e->dtype = static_cast<int>(v); break;
      case 2:
        WalkProto(b, [&](int f2, int, uint64_t, Slice dim) {
          if (f2 != 2) return;
          int64_t sz = 0;
          WalkProto(dim, [&](int f3, int, uint64_t v3, Slice) {
            if (f3 == 1) sz = static_cast<int64_t>(v3);
          });
          e->shape.push_back(sz);
        });
        break;
      case 3: e->shard = static_cast<int64_t>(v); break;
      case 4: e->offset = static_cast<int64_t>(v); break;
      case 5: e->size = static_cast<int64_t>(v); break;
      case 6: e->crc = static_cast<uint32_t>(v); break;
      default: break;
    }
  });
}

}  // namespace

uint32_t Crc32c(const uint8_t* p, size_t n) {
  const auto& T = Table().t;
  uint32_t c = 0xffffffffu;
  while (n >= 8) {
    uint32_t lo, hi;
    std::memcpy(&lo, p, 4);
    std::memcpy(&hi, p + 4, 4);
    lo ^= c;
    c = T[7][lo & 0xff] ^ T[6][(lo >> 8) & 0xff] ^ T[5][(lo >> 16) & 0xff] ^ T[4][lo >> 24] ^
        T[3][hi & 0xff] ^ T[2][(hi >> 8) & 0xff] ^ T[1][(hi >> 16) & 0xff] ^ T[0][hi >> 24];
    p += 8;
    n -= 8;
  }
  while (n--) c = T[0][(c ^ *p++) & 0xff] ^ (c >> 8);
  return c ^ 0xffffffffu;
}

uint32_t MaskCrc(uint32_t crc) { return ((crc >> 15) | (crc << 17)) + 0xa282ead8u; }

BundleError ReadBundle(const std::string& prefix, TensorMap* out, std::string* err) {
  std::vector<uint8_t> index, data;
  if (!ReadFile(prefix + ".index", &index)) {
    *err = "cannot read " + prefix + ".index";
    return BundleError::kIo;
  }
  if (index.size() < 48) {
    *err = "index file too short";
    return BundleError::kFormat;
  }
  const uint8_t* footer = index.data() + index.size() - 48;
  uint64_t magic;
  std::memcpy(&magic, footer + 40, 8);
  if (magic != kTableMagic) {
    *err = "bad SSTable magic in " + prefix + ".index";
    return BundleError::kFormat;
  }
  Cursor fc{footer, footer + 40};
  fc.Varint();
  fc.Varint();
  uint64_t idx_off = fc.Varint(), idx_size = fc.Varint();
  if (!fc.ok) {
    *err = "corrupt SSTable footer";
    return BundleError::kFormat;
  }
  Slice index_block;
  if (!GetBlock(index, idx_off, idx_size, &index_block, err)) return BundleError::kFormat;

  std::map<std::string, Entry> entries;
  int64_t num_shards = 1, endianness = 0;
  bool ok = ForEachEntry(
      index_block,
      [&](const std::string&, Slice handle) {
        Cursor hc{handle.p, handle.p + handle.n};
        uint64_t off = hc.Varint(), size = hc.Varint();
        Slice block;
        if (!hc.ok || !GetBlock(index, off, size, &block, err)) return false;
        return ForEachEntry(
            block,
            [&](const std::string& key, Slice val) {
              if (key.empty()) {
                num_shards = 0;
                return WalkProto(val, [&](int f, int, uint64_t v, Slice) {
                  if (f == 1) num_shards = static_cast<int64_t>(v);
                  if (f == 2) endianness = static_cast<int64_t>(v);
                });
              }
              Entry e;
              if (!ParseEntry(val, &e)) {
                *err = "corrupt BundleEntryProto for " + key;
                return false;
              }
              entries[key] = e;
              return true;
            },
            err);
      },
      err);
  if (!ok) {
    if (err->empty()) *err = "corrupt index";
    return BundleError::kFormat;
  }
  if (num_shards != 1 || endianness != 0) {
    *err = "only single-shard little-endian bundles are supported";
    return BundleError::kFormat;
  }
  if (!ReadFile(prefix + ".data-00000-of-00001", &data)) {
    *err = "cannot read " + prefix + ".data-00000-of-00001";
    return BundleError::kIo;
  }
  out->clear();
  for (auto& kv : entries) {
    const Entry& e = kv.second;
    if (e.dtype != 1) continue;  // only DT_FLOAT variables matter on this path
    Tensor t;
    t.shape = e.shape;
    // untrusted header: reject negative / overflowing dims, offsets and sizes before any arithmetic on them
    bool dims_ok = true;
    uint64_t numel = 1;
    for (int64_t d : e.shape) {
      if (d < 0 || (d > 0 && numel > (static_cast<uint64_t>(1) << 40) / static_cast<uint64_t>(d))) {
        dims_ok = false;
        break;
      }
      numel *= static_cast<uint64_t>(d);
    }
    const uint64_t dsz = data.size();
    if (!dims_ok || e.shard != 0 || e.offset < 0 || e.size < 0 || static_cast<uint64_t>(e.size) != numel * 4 ||
        static_cast<uint64_t>(e.size) > dsz || static_cast<uint64_t>(e.offset) > dsz - static_cast<uint64_t>(e.size)) {
      *err = "tensor " + kv.first + ": bad shape/offset/size";
      return BundleError::kFormat;
    }
    if (MaskCrc(Crc32c(&data[e.offset], static_cast<size_t>(e.size))) != e.crc) {
      *err = "tensor " + kv.first + ": CRC32C mismatch";
      return BundleError::kFormat;
    }
    t.data.resize(static_cast<size_t>(t.numel()));
    std::memcpy(t.data.data(), &data[e.offset], static_cast<size_t>(e.size));
    (*out)[kv.first] = std::move(t);
  }
  return BundleError::kOk;
}

}  // namespace rn
