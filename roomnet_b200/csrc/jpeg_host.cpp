// JPEG marker parser + Huffman decoder (host side of the JPEG front end, see jpeg_host.h).
// Written against ITU-T T.81 (baseline sequential DCT, Annex B marker syntax, Annex F Huffman decoding); the geometry
// rules (which blocks of an interleaved MCU carry image data, downsampled sizes) are the ones the reference's decoder
// applies - cv2.imread -> libjpeg-turbo - so that the coefficient arrays line up with its inverse-DCT stage.
#include "jpeg_host.h"

#include <cstring>

namespace rn {
namespace {

const uint8_t kZigzag[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                             41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                             30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

inline int CeilDiv(int a, int b) { return (a + b - 1) / b; }
inline unsigned Be16(const uint8_t* p) { return (static_cast<unsigned>(p[0]) << 8) | p[1]; }

// EXIF orientation from an APP1 payload; returns false when the segment claims to be EXIF but cannot be read
bool ExifOrientation(const uint8_t* p, size_t n, int* orientation) {
  if (n < 6 || std::memcmp(p, "Exif\0\0", 6) != 0) return true;  // XMP or something else: no orientation here
  p += 6;
  n -= 6;
  if (n < 8) return false;
  bool le;
  if (p[0] == 'I' && p[1] == 'I') le = true;
  else if (p[0] == 'M' && p[1] == 'M') le = false;
  else return false;
  auto u16 = [&](size_t o) -> unsigned { return le ? (p[o] | (p[o + 1] << 8)) : ((p[o] << 8) | p[o + 1]); };
  auto u32 = [&](size_t o) -> size_t {
    return le ? (static_cast<size_t>(p[o]) | (static_cast<size_t>(p[o + 1]) << 8) | (static_cast<size_t>(p[o + 2]) << 16) |
                 (static_cast<size_t>(p[o + 3]) << 24))
              : ((static_cast<size_t>(p[o]) << 24) | (static_cast<size_t>(p[o + 1]) << 16) |
                 (static_cast<size_t>(p[o + 2]) << 8) | static_cast<size_t>(p[o + 3]));
  };
  if (u16(2) != 42) return false;
  const size_t ifd = u32(4);
  if (ifd > n || n - ifd < 2) return false;
  const unsigned count = u16(ifd);
  if ((n - ifd - 2) / 12 < count) return false;
  for (unsigned i = 0; i < count; ++i) {
    const size_t e = ifd + 2 + 12 * static_cast<size_t>(i);
    if (u16(e) == 0x0112) {
      if (u16(e + 2) != 3 || u32(e + 4) != 1) return false;
      const unsigned o = u16(e + 8);
      *orientation = (o >= 1 && o <= 8) ? static_cast<int>(o) : 1;
      return true;
    }
  }
  return true;
}

struct Huff {
  bool defined = false;
  uint16_t lookup[512];  // (length << 8) | symbol for codes of <= 9 bits, 0 = longer code
  int maxcode[18];       // largest code of each length (-1 = none), [17] = sentinel
  int mincode[17];
  int valptr[17];
  uint8_t vals[256];
  // AC fast path, indexed by the next kFastBits bits: code and value bits together when both fit.
  // bits 0-7 = bits consumed (0 = take the slow path), 8-14 = zero run to skip (64 = end of block), 15 = nothing to
  // store (ZRL), 16-31 = the coefficient value (signed)
  int32_t fast[1 << 11];
};
constexpr int kFastBits = 11;

bool BuildHuff(const uint8_t* counts, const uint8_t* symbols, int nsym, Huff* t) {
  std::memset(t->lookup, 0, sizeof(t->lookup));
  std::memcpy(t->vals, symbols, nsym);
  int code = 0, k = 0;
  for (int len = 1; len <= 16; ++len) {
    t->valptr[len] = k;
    t->mincode[len] = code;
    for (int i = 0; i < counts[len - 1]; ++i, ++k, ++code) {
      if (code >= (1 << len)) return false;  // over-subscribed
      if (len <= 9) {
        const int first = code << (9 - len);
        for (int j = 0; j < (1 << (9 - len)); ++j) t->lookup[first + j] = static_cast<uint16_t>((len << 8) | symbols[k]);
      }
    }
    t->maxcode[len] = counts[len - 1] ? code - 1 : -1;
    code <<= 1;
  }
  t->maxcode[17] = 0x7fffffff;
  t->defined = true;
  // fast table (meaningful for AC tables; harmless for DC ones, which do not use it)
  std::memset(t->fast, 0, sizeof(t->fast));
  code = 0;
  k = 0;
  for (int len = 1; len <= kFastBits; ++len) {
    for (int i = 0; i < counts[len - 1]; ++i, ++k, ++code) {
      const int rs = symbols[k], r = rs >> 4, sz = rs & 15;
      if (len + sz > kFastBits) continue;
      const int free_bits = kFastBits - len - sz;
      for (int v = 0; v < (1 << sz); ++v) {
        int32_t e;
        if (sz == 0) {
          e = (r == 15) ? ((16 << 8) | (1 << 15) | len) : (r == 0 ? ((64 << 8) | (1 << 15) | len) : 0);
          if (e == 0) continue;  // run/size with size 0 other than EOB / ZRL: undefined in baseline, slow path rejects it
        } else {
          const int val = v < (1 << (sz - 1)) ? v - (1 << sz) + 1 : v;
          e = static_cast<int32_t>(static_cast<uint32_t>(val) << 16) | (r << 8) | (len + sz);
        }
        const int first = ((code << sz) | v) << free_bits;
        for (int j = 0; j < (1 << free_bits); ++j) t->fast[first + j] = e;
      }
    }
    code <<= 1;
  }
  return true;
}

// MSB-first bit reader over an entropy-coded segment: removes the 0x00 stuffed after 0xFF, stops at the first marker
// and feeds zero bits from there on (counted, so that a scan that consumed them is reported as damaged).
struct BitReader {
  const uint8_t* p;
  const uint8_t* end;
  uint64_t buf = 0;
  int cnt = 0;        // valid bits at the top of buf
  int pad_bits = 0;   // zero bits appended after the data ran out
  bool at_marker = false;

  // Refill.  Common case: the next eight bytes hold no 0xFF, so they go in with one load; whole bytes are counted,
  // the bits of the partial byte below them are OR-ed in again (identically) by the next refill.
  inline void Fill() {
    if (!at_marker && end - p >= 8) {
      uint64_t raw;
      std::memcpy(&raw, p, 8);
      const uint64_t inv = ~raw;  // a 0xFF byte in raw is a zero byte here
      if (!((inv - 0x0101010101010101ull) & ~inv & 0x8080808080808080ull)) {
        buf |= __builtin_bswap64(raw) >> cnt;
        const int adv = (63 - cnt) >> 3;
        p += adv;
        cnt += adv * 8;
        return;
      }
    }
    FillSlow();
  }
  void FillSlow() {
    buf &= cnt ? ~0ull << (64 - cnt) : 0ull;  // drop the partial-byte bits of a fast refill: bytes are re-read below
    while (cnt <= 56) {
      unsigned b = 0;
      if (!at_marker && p < end) {
        b = *p;
        if (b == 0xFF) {
          if (p + 1 < end && p[1] == 0x00) {
            p += 2;
          } else {
            at_marker = true;  // p stays on the 0xFF of the marker
            b = 0;
            pad_bits += 8;
          }
        } else {
          ++p;
        }
      } else {
        at_marker = true;
        pad_bits += 8;
      }
      buf |= static_cast<uint64_t>(b) << (56 - cnt);
      cnt += 8;
    }
  }
  inline unsigned Peek(int n) const { return static_cast<unsigned>(buf >> (64 - n)); }
  inline void Skip(int n) {
    buf <<= n;
    cnt -= n;
  }
  inline bool Overrun() const { return cnt < pad_bits; }
  void Reset() {
    buf = 0;
    cnt = 0;
    pad_bits = 0;
    at_marker = false;
  }
};

inline int DecodeSymbol(BitReader& br, const Huff& t) {
  const unsigned e = t.lookup[br.Peek(9)];
  if (e) {
    br.Skip(e >> 8);
    return e & 0xff;
  }
  const unsigned code16 = br.Peek(16);
  for (int len = 10; len <= 16; ++len) {
    const int code = static_cast<int>(code16 >> (16 - len));
    if (code <= t.maxcode[len]) {
      br.Skip(len);
      return t.vals[t.valptr[len] + code - t.mincode[len]];
    }
  }
  return -1;
}

inline int Extend(unsigned v, int s) { return v < (1u << (s - 1)) ? static_cast<int>(v) - (1 << s) + 1 : static_cast<int>(v); }

// one 8x8 block (T.81 F.2.2); `out` may be a scratch block for the dummy blocks of an edge MCU
bool DecodeBlock(BitReader& br, const Huff& dc, const Huff& ac, int* pred, int16_t* out) {
  if (br.cnt < 32) br.Fill();
  int s = DecodeSymbol(br, dc);
  if (s < 0 || s > 15) return false;
  if (s) {
    const unsigned v = br.Peek(s);
    br.Skip(s);
    *pred += Extend(v, s);
  }
  out[0] = static_cast<int16_t>(*pred);
  for (int k = 1; k < 64;) {
    if (br.cnt < 32) br.Fill();
    const int32_t e = ac.fast[br.Peek(kFastBits)];
    if (e & 0xff) {
      br.Skip(e & 0xff);
      k += (e >> 8) & 0x7f;
      if (e & (1 << 15)) continue;  // ZRL, or end of block (k is now >= 64)
      if (k > 63) return false;
      out[kZigzag[k]] = static_cast<int16_t>(e >> 16);
      ++k;
      continue;
    }
    const int rs = DecodeSymbol(br, ac);
    if (rs < 0) return false;
    const int r = rs >> 4;
    s = rs & 15;
    if (s == 0) {
      if (r != 15) break;  // end of block
      k += 16;
      continue;
    }
    k += r;
    if (k > 63) return false;
    const unsigned v = br.Peek(s);
    br.Skip(s);
    out[kZigzag[k]] = static_cast<int16_t>(Extend(v, s));
    ++k;
  }
  return true;
}

struct Segment {  // one marker segment with a length field
  int marker;
  const uint8_t* payload;
  size_t len;
};

// advances *pos to the next marker and returns it (0 at the end of the data)
int NextMarker(const uint8_t* data, size_t size, size_t* pos) {
  size_t i = *pos;
  while (i < size && data[i] != 0xFF) ++i;  // tolerate garbage between segments like the reference decoder
  while (i < size && data[i] == 0xFF) ++i;  // fill bytes
  if (i >= size) return 0;
  *pos = i + 1;
  return data[i];
}

bool ReadSegment(const uint8_t* data, size_t size, size_t* pos, Segment* s) {
  if (size - *pos < 2) return false;
  const size_t len = Be16(data + *pos);
  if (len < 2 || len > size - *pos) return false;
  s->payload = data + *pos + 2;
  s->len = len - 2;
  *pos += len;
  return true;
}

bool ParseDqt(const Segment& s, JpegInfo* info) {
  const uint8_t* p = s.payload;
  size_t n = s.len;
  while (n > 0) {
    const int pq = p[0] >> 4, tq = p[0] & 15;
    if (tq > 3 || pq > 1) return false;
    const size_t need = 1 + 64 * (pq ? 2 : 1);
    if (n < need) return false;
    for (int i = 0; i < 64; ++i) {
      const unsigned q = pq ? Be16(p + 1 + 2 * i) : p[1 + i];
      info->quant[tq][kZigzag[i]] = static_cast<uint16_t>(q);
    }
    info->have_quant[tq] = true;
    p += need;
    n -= need;
  }
  return true;
}

bool ParseDht(const Segment& s, Huff dc[4], Huff ac[4]) {
  const uint8_t* p = s.payload;
  size_t n = s.len;
  while (n > 0) {
    if (n < 17) return false;
    const int tc = p[0] >> 4, th = p[0] & 15;
    if (tc > 1 || th > 3) return false;
    int nsym = 0;
    for (int i = 0; i < 16; ++i) nsym += p[1 + i];
    if (nsym > 256 || n < static_cast<size_t>(17 + nsym)) return false;
    if (!BuildHuff(p + 1, p + 17, nsym, tc ? &ac[th] : &dc[th])) return false;
    p += 17 + nsym;
    n -= 17 + nsym;
  }
  return true;
}

}  // namespace

JpegStatus JpegParseHeader(const uint8_t* data, size_t size, JpegInfo* info) {
  *info = JpegInfo{};
  if (!data || size < 4 || data[0] != 0xFF || data[1] != 0xD8) return kJpegCorrupt;
  size_t pos = 2;
  bool have_sof = false, saw_jfif = false, saw_adobe = false, saw_exif = false;
  int adobe_transform = 0;
  for (;;) {
    const int m = NextMarker(data, size, &pos);
    if (m == 0 || m == 0xD9) return kJpegCorrupt;  // no scan
    if (m == 0x01 || (m >= 0xD0 && m <= 0xD7)) continue;  // TEM / stray RSTn: no payload
    Segment s;
    if (!ReadSegment(data, size, &pos, &s)) return kJpegCorrupt;
    if (m == 0xC0 || m == 0xC1) {
      if (have_sof || s.len < 6) return kJpegCorrupt;
      if (s.payload[0] != 8) return kJpegUnsupported;  // 12-bit samples
      info->height = static_cast<int>(Be16(s.payload + 1));
      info->width = static_cast<int>(Be16(s.payload + 3));
      info->ncomp = s.payload[5];
      if (info->height == 0 || info->width == 0) return kJpegUnsupported;  // DNL-defined height
      if (info->ncomp != 1 && info->ncomp != 3) return kJpegUnsupported;   // CMYK / YCCK / two-channel
      if (s.len < static_cast<size_t>(6 + 3 * info->ncomp)) return kJpegCorrupt;
      for (int c = 0; c < info->ncomp; ++c) {
        const uint8_t* q = s.payload + 6 + 3 * c;
        info->comp[c].id = q[0];
        info->comp[c].h = q[1] >> 4;
        info->comp[c].v = q[1] & 15;
        info->comp[c].tq = q[2];
        if (info->comp[c].h < 1 || info->comp[c].h > 4 || info->comp[c].v < 1 || info->comp[c].v > 4 || q[2] > 3)
          return kJpegCorrupt;
      }
      have_sof = true;
    } else if ((m >= 0xC2 && m <= 0xCF) && m != 0xC4 && m != 0xC8) {
      return kJpegUnsupported;  // progressive, lossless, differential, arithmetic (incl. DAC)
    } else if (m == 0xDB) {
      if (!ParseDqt(s, info)) return kJpegCorrupt;
    } else if (m == 0xDD) {
      if (s.len < 2) return kJpegCorrupt;
      info->restart_interval = static_cast<int>(Be16(s.payload));
    } else if (m == 0xE0) {
      if (s.len >= 5 && std::memcmp(s.payload, "JFIF\0", 5) == 0) saw_jfif = true;
    } else if (m == 0xEE) {
      if (s.len >= 12 && std::memcmp(s.payload, "Adobe", 5) == 0) {
        saw_adobe = true;
        adobe_transform = s.payload[11];
      }
    } else if (m == 0xE1) {
      if (!saw_exif && s.len >= 6 && std::memcmp(s.payload, "Exif\0\0", 6) == 0) {
        saw_exif = true;
        if (!ExifOrientation(s.payload, s.len, &info->orientation)) return kJpegUnsupported;
      }
    } else if (m == 0xDA) {
      if (!have_sof) return kJpegCorrupt;
      info->sos_offset = pos - (s.len + 2) - 2;  // the 0xFF of the SOS marker
      break;
    }
    // DHT and everything else: handled (or skipped) by the decode pass
  }
  // colour space as the reference decoder deduces it (JFIF => YCbCr; Adobe transform; else by component ids)
  if (info->ncomp == 3) {
    bool ycc = true;
    if (saw_jfif) ycc = true;
    else if (saw_adobe) ycc = adobe_transform != 0;
    else if (info->comp[0].id == 'R' && info->comp[1].id == 'G' && info->comp[2].id == 'B') ycc = false;
    if (!ycc) return kJpegUnsupported;
    if (info->comp[1].h != 1 || info->comp[1].v != 1 || info->comp[2].h != 1 || info->comp[2].v != 1) return kJpegUnsupported;
    const int h = info->comp[0].h, v = info->comp[0].v;
    if (!((h == 1 && v == 1) || (h == 2 && v == 1) || (h == 2 && v == 2))) return kJpegUnsupported;
  } else {
    info->comp[0].h = info->comp[0].v = 1;  // a single component is never subsampled against itself
  }
  if (info->width < 16 || info->height < 16) return kJpegUnsupported;  // the decoder's tiny-image special cases
  info->hmax = info->comp[0].h;
  info->vmax = info->comp[0].v;
  size_t off = 0;
  for (int c = 0; c < info->ncomp; ++c) {
    JpegComponent& k = info->comp[c];
    if (!info->have_quant[k.tq]) {
      // a table defined after the first SOS is legal but not something an encoder of photographs emits
      return kJpegUnsupported;
    }
    k.dw = CeilDiv(info->width * k.h, info->hmax);
    k.dh = CeilDiv(info->height * k.v, info->vmax);
    k.wblocks = CeilDiv(k.dw, 8);
    k.hblocks = CeilDiv(k.dh, 8);
    k.coef_offset = off;
    off += static_cast<size_t>(k.wblocks) * k.hblocks * 64;
  }
  info->coef_count = off;
  return kJpegOk;
}

JpegStatus JpegDecodeCoefficients(const uint8_t* data, size_t size, const JpegInfo& info, int16_t* coefs) {
  std::memset(coefs, 0, info.coef_count * sizeof(int16_t));
  Huff dc[4], ac[4];
  bool decoded[3] = {false, false, false};
  int restart_interval = 0;
  size_t pos = 2;
  for (;;) {
    const int m = NextMarker(data, size, &pos);
    if (m == 0) return kJpegCorrupt;
    if (m == 0xD9) break;
    if (m == 0x01 || (m >= 0xD0 && m <= 0xD7)) continue;
    Segment s;
    if (!ReadSegment(data, size, &pos, &s)) return kJpegCorrupt;
    if (m == 0xC4) {
      if (!ParseDht(s, dc, ac)) return kJpegCorrupt;
    } else if (m == 0xDD) {
      if (s.len < 2) return kJpegCorrupt;
      restart_interval = static_cast<int>(Be16(s.payload));
    } else if (m == 0xDB) {
      if (decoded[0] || decoded[1] || decoded[2]) return kJpegUnsupported;  // tables redefined between scans
    } else if (m == 0xDA) {
      if (s.len < 1) return kJpegCorrupt;
      const int ns = s.payload[0];
      if (ns < 1 || ns > info.ncomp || s.len < static_cast<size_t>(4 + 2 * ns)) return kJpegCorrupt;
      int ci[3], td[3], ta[3];
      for (int i = 0; i < ns; ++i) {
        const int id = s.payload[1 + 2 * i];
        ci[i] = -1;
        for (int c = 0; c < info.ncomp; ++c)
          if (info.comp[c].id == id) ci[i] = c;
        if (ci[i] < 0 || decoded[ci[i]]) return kJpegCorrupt;
        for (int j = 0; j < i; ++j)
          if (ci[j] == ci[i]) return kJpegCorrupt;
        td[i] = s.payload[2 + 2 * i] >> 4;
        ta[i] = s.payload[2 + 2 * i] & 15;
        if (td[i] > 3 || ta[i] > 3 || !dc[td[i]].defined || !ac[ta[i]].defined) return kJpegUnsupported;  // e.g. MJPEG frames without DHT
      }
      const uint8_t* q = s.payload + 1 + 2 * ns;
      if (q[0] != 0 || q[1] != 63 || q[2] != 0) return kJpegUnsupported;  // spectral selection / successive approximation
      // ---- entropy-coded segment ----
      BitReader br;
      br.p = data + pos;
      br.end = data + size;
      int pred[3] = {0, 0, 0};
      int16_t dummy[64];
      const bool interleaved = ns > 1;
      const int mcus_x = interleaved ? CeilDiv(info.width, 8 * info.hmax) : info.comp[ci[0]].wblocks;
      const int mcus_y = interleaved ? CeilDiv(info.height, 8 * info.vmax) : info.comp[ci[0]].hblocks;
      int until_restart = restart_interval ? restart_interval : -1;
      int next_rst = 0;
      for (int my = 0; my < mcus_y; ++my) {
        for (int mx = 0; mx < mcus_x; ++mx) {
          if (until_restart == 0) {
            // byte-align, expect RSTn
            if (br.Overrun()) return kJpegCorrupt;
            if (!br.at_marker && (br.cnt - br.pad_bits) >= 8) return kJpegCorrupt;  // whole data bytes before the marker
            const uint8_t* p = br.p;
            if (!br.at_marker) {
              while (p < br.end && *p != 0xFF) ++p;
            }
            while (p + 1 < br.end && p[0] == 0xFF && p[1] == 0xFF) ++p;
            if (p + 1 >= br.end || p[0] != 0xFF || p[1] != 0xD0 + next_rst) return kJpegCorrupt;
            br.p = p + 2;
            br.Reset();
            next_rst = (next_rst + 1) & 7;
            pred[0] = pred[1] = pred[2] = 0;
            until_restart = restart_interval;
          }
          for (int i = 0; i < ns; ++i) {
            const JpegComponent& k = info.comp[ci[i]];
            int16_t* base = coefs + k.coef_offset;
            const int bh = interleaved ? k.h : 1, bv = interleaved ? k.v : 1;
            for (int vv = 0; vv < bv; ++vv) {
              for (int hh = 0; hh < bh; ++hh) {
                const int bx = mx * bh + hh, by = my * bv + vv;
                int16_t* out = dummy;
                if (bx < k.wblocks && by < k.hblocks) {
                  out = base + (static_cast<size_t>(by) * k.wblocks + bx) * 64;
                }
                if (!DecodeBlock(br, dc[td[i]], ac[ta[i]], &pred[i], out)) return kJpegCorrupt;
              }
            }
          }
          if (until_restart > 0) --until_restart;
        }
      }
      if (br.Overrun()) return kJpegCorrupt;  // the scan ended inside the zero padding: truncated file
      // continue at the marker that ended the segment
      if (br.at_marker) {
        pos = static_cast<size_t>(br.p - data);
      } else {
        const uint8_t* p = br.p;
        while (p + 1 < br.end && !(p[0] == 0xFF && p[1] != 0x00 && !(p[1] >= 0xD0 && p[1] <= 0xD7))) ++p;
        pos = static_cast<size_t>(p - data);
      }
      for (int i = 0; i < ns; ++i) decoded[ci[i]] = true;
    }
  }
  for (int c = 0; c < info.ncomp; ++c)
    if (!decoded[c]) return kJpegCorrupt;
  return kJpegOk;
}

namespace {
bool BuildDevHuff(const uint8_t* counts, const uint8_t* symbols, int nsym, DevHuffTable* t) {
  std::memset(t, 0, sizeof(*t));
  std::memcpy(t->vals, symbols, nsym);
  int code = 0, k = 0;
  for (int len = 1; len <= 16; ++len) {
    t->valoff[len] = k - code;
    for (int i = 0; i < counts[len - 1]; ++i, ++k, ++code) {
      if (code >= (1 << len)) return false;
      if (len <= 10) {
        const int first = code << (10 - len);
        for (int j = 0; j < (1 << (10 - len)); ++j) t->fast[first + j] = static_cast<uint16_t>((len << 8) | symbols[k]);
      }
    }
    t->maxcode[len] = counts[len - 1] ? code - 1 : -1;
    code <<= 1;
  }
  t->maxcode[17] = 0x7fffffff;
  return true;
}
}  // namespace

size_t JpegStreamCapacity(const JpegInfo& info, size_t size) {
  const size_t mcus = static_cast<size_t>(CeilDiv(info.width, 8 * info.hmax)) * CeilDiv(info.height, 8 * info.vmax);
  const size_t segs = info.restart_interval ? (mcus + info.restart_interval - 1) / info.restart_interval : 1;
  return (size / kSubseqBytes + segs + 2) * kSubseqBytes;
}

JpegStatus JpegPrepareScan(const uint8_t* data, size_t size, const JpegInfo& info, JpegScanPlan* plan, uint8_t* stream,
                           size_t stream_capacity, int32_t* sub_seg) {
  // ---- tables and the scan header ----
  struct RawHuff {
    bool defined = false;
    uint8_t counts[16];
    uint8_t symbols[256];
    int nsym = 0;
  } dc[4], ac[4];
  size_t pos = 2;
  int restart_interval = 0;
  const uint8_t* ecs = nullptr;
  int td[3] = {0, 0, 0}, ta[3] = {0, 0, 0};
  for (;;) {
    const int m = NextMarker(data, size, &pos);
    if (m == 0 || m == 0xD9) return kJpegCorrupt;
    if (m == 0x01 || (m >= 0xD0 && m <= 0xD7)) continue;
    Segment s;
    if (!ReadSegment(data, size, &pos, &s)) return kJpegCorrupt;
    if (m == 0xC4) {
      const uint8_t* p = s.payload;
      size_t n = s.len;
      while (n > 0) {
        if (n < 17) return kJpegCorrupt;
        const int tc = p[0] >> 4, th = p[0] & 15;
        if (tc > 1 || th > 3) return kJpegCorrupt;
        int nsym = 0;
        for (int i = 0; i < 16; ++i) nsym += p[1 + i];
        if (nsym > 256 || n < static_cast<size_t>(17 + nsym)) return kJpegCorrupt;
        RawHuff& r = tc ? ac[th] : dc[th];
        r.defined = true;
        std::memcpy(r.counts, p + 1, 16);
        std::memcpy(r.symbols, p + 17, nsym);
        r.nsym = nsym;
        p += 17 + nsym;
        n -= 17 + nsym;
      }
    } else if (m == 0xDD) {
      if (s.len < 2) return kJpegCorrupt;
      restart_interval = static_cast<int>(Be16(s.payload));
    } else if (m == 0xDA) {
      const int ns = s.len >= 1 ? s.payload[0] : 0;
      if (ns != info.ncomp || s.len < static_cast<size_t>(4 + 2 * ns)) return kJpegUnsupported;  // several scans
      for (int i = 0; i < ns; ++i) {
        if (s.payload[1 + 2 * i] != info.comp[i].id) return kJpegUnsupported;  // components out of frame order
        td[i] = s.payload[2 + 2 * i] >> 4;
        ta[i] = s.payload[2 + 2 * i] & 15;
        if (td[i] > 3 || ta[i] > 3 || !dc[td[i]].defined || !ac[ta[i]].defined) return kJpegUnsupported;
      }
      const uint8_t* q = s.payload + 1 + 2 * ns;
      if (q[0] != 0 || q[1] != 63 || q[2] != 0) return kJpegUnsupported;
      ecs = data + pos;
      break;
    }
  }
  for (int c = 0; c < info.ncomp; ++c) {
    if (!BuildDevHuff(dc[td[c]].counts, dc[td[c]].symbols, dc[td[c]].nsym, &plan->tab[2 * c]) ||
        !BuildDevHuff(ac[ta[c]].counts, ac[ta[c]].symbols, ac[ta[c]].nsym, &plan->tab[2 * c + 1]))
      return kJpegCorrupt;
  }
  // ---- MCU structure ----
  plan->bpm = 0;
  for (int c = 0; c < info.ncomp; ++c)
    for (int vv = 0; vv < info.comp[c].v; ++vv)
      for (int hh = 0; hh < info.comp[c].h; ++hh) {
        plan->blk_comp[plan->bpm] = static_cast<uint8_t>(c);
        plan->blk_hh[plan->bpm] = static_cast<uint8_t>(hh);
        plan->blk_vv[plan->bpm] = static_cast<uint8_t>(vv);
        ++plan->bpm;
      }
  if (info.ncomp == 1) {  // a single-component scan is not interleaved: one block per MCU, no dummy blocks
    plan->mcus_x = info.comp[0].wblocks;
    plan->mcus_y = info.comp[0].hblocks;
  } else {
    plan->mcus_x = CeilDiv(info.width, 8 * info.hmax);
    plan->mcus_y = CeilDiv(info.height, 8 * info.vmax);
  }
  const size_t mcus = static_cast<size_t>(plan->mcus_x) * plan->mcus_y;
  if (mcus * plan->bpm >= (size_t{1} << 31)) return kJpegUnsupported;
  plan->total_blocks = static_cast<uint32_t>(mcus * plan->bpm);
  plan->seg_blocks = restart_interval ? static_cast<uint32_t>(restart_interval) * plan->bpm : plan->total_blocks;
  plan->n_seg = restart_interval ? static_cast<int>((mcus + restart_interval - 1) / restart_interval) : 1;
  // ---- unstuff; restart segments start on subsequence boundaries ----
  const uint8_t* p = ecs;
  const uint8_t* end = data + size;
  size_t o = 0;
  int seg = 0;
  size_t seg_start = 0;
  bool eoi = false;
  while (p < end) {
    const uint8_t* ff = static_cast<const uint8_t*>(std::memchr(p, 0xFF, static_cast<size_t>(end - p)));
    const size_t run = static_cast<size_t>((ff ? ff : end) - p);
    if (o + run + 2 * kSubseqBytes > stream_capacity) return kJpegCorrupt;
    std::memcpy(stream + o, p, run);
    o += run;
    if (!ff) {
      p = end;
      break;
    }
    p = ff + 1;
    if (p >= end) break;
    const int m = *p;
    if (m == 0x00) {
      stream[o++] = 0xFF;
      ++p;
    } else if (m == 0xFF) {
      // fill byte: the next 0xFF decides
    } else if (m >= 0xD0 && m <= 0xD7) {
      if (!restart_interval || m != 0xD0 + (seg & 7) || seg + 1 >= plan->n_seg) return kJpegCorrupt;
      ++p;
      const size_t padded = (o + kSubseqBytes - 1) / kSubseqBytes * kSubseqBytes;
      std::memset(stream + o, 0, padded - o);
      o = padded;
      for (size_t i = seg_start / kSubseqBytes; i < o / kSubseqBytes; ++i) sub_seg[i] = seg;
      if (o == seg_start) return kJpegCorrupt;  // an empty restart segment
      seg_start = o;
      ++seg;
    } else {
      eoi = (m == 0xD9);
      break;  // a marker other than RSTn ends the scan
    }
  }
  if (!eoi) return kJpegUnsupported;  // truncated, or more segments follow (another scan): the host decoder decides
  if (seg + 1 != plan->n_seg) return kJpegCorrupt;
  const size_t padded = (o + kSubseqBytes - 1) / kSubseqBytes * kSubseqBytes;
  if (padded + kSubseqBytes > stream_capacity) return kJpegCorrupt;
  std::memset(stream + o, 0, padded - o);
  o = padded;
  if (o == seg_start) return kJpegCorrupt;
  for (size_t i = seg_start / kSubseqBytes; i < o / kSubseqBytes; ++i) sub_seg[i] = seg;
  plan->stream_bytes = o;
  return kJpegOk;
}

}  // namespace rn
