// Huffman (entropy) decoding of baseline JPEG scans on the device - the serial half of cv2.imread (infer.py:81).
//
// A Huffman stream has no random access, but a decoder started at an arbitrary bit re-synchronises with the true
// symbol boundaries after a few symbols.  The stream (byte stuffing removed by the host while it gathers the files) is
// cut into subsequences of 1024 bits, one thread each:
//   1. every thread decodes its subsequence from a guessed state and remembers the state it ends in
//      (bit position, block of the MCU, zigzag index);
//   2. every thread whose predecessor ended in a state other than the one it started from decodes again from there -
//      repeated (inside a CUDA block through shared memory, between blocks by relaunching) until nothing changes.  The
//      first subsequence of a restart segment starts from a known state, so the fixed point is the sequential decode;
//   3. a segmented prefix sum over the blocks completed per subsequence gives every thread its absolute block number;
//   4. a last pass decodes once more and writes the coefficients where the inverse-DCT kernel expects them; the DC
//      differences go to a scan-order array and are summed per component (predictor reset at restart markers) by a
//      segmented scan.
// Integer / bit work throughout; results are bit-identical to the host decoder (jpeg_host.cpp) by construction and the
// tests compare the decoded image with cv2's.
#include "kernels.h"

namespace rn {
namespace {

constexpr int kSubBits = kSubseqBytes * 8;
constexpr int kHuffThreads = 256;  // subsequences per CUDA block
constexpr int kMaxInner = 64;      // re-synchronisation sweeps inside a block per launch
constexpr int kMaxRounds = 16;     // launches before the batch is handed to the host decoder
constexpr int kGuessBits = 256;    // how much of a subsequence the first (guessed) decode covers

__device__ __constant__ unsigned char kZigzagDev[80] = {
    0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13,
    6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31,
    39, 46, 53, 60, 61, 54, 47, 55, 62, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63};

struct SmemHuff {
  unsigned char zigzag[80];  // (constant memory serialises lanes that index it differently; shared memory does not)
  DevHuffTable tab[6];
  unsigned words[kHuffThreads * kSubseqBytes / 4 + 8];  // the block's part of the stream (+ 32 bytes of look-ahead)
};

// decoder state between two symbols
__device__ __forceinline__ unsigned long long pack_state(unsigned bit, unsigned blk, unsigned k) {
  return static_cast<unsigned long long>(bit) | (static_cast<unsigned long long>(blk) << 32) |
         (static_cast<unsigned long long>(k) << 40);
}

struct WriteCtx {
  const HuffFileDesc* f;
  int16_t* coefs;
  int16_t* dcdiff;
  unsigned block;      // absolute block number (scan order) of the block being decoded
  unsigned seg_end;    // first block of the next restart segment
  int16_t* dst;        // coefficient block being written (nullptr = dummy edge block)
  int* error;
  unsigned mx, my;     // MCU column / row of `block` (kept in step with it: no divisions per block)
};

// destination of block j of MCU (mx, my)
__device__ __forceinline__ int16_t* mcu_block_dst(const HuffFileDesc& f, int16_t* coefs, unsigned mx, unsigned my, unsigned j,
                                                  unsigned block) {
  if (f.ncomp == 1) return coefs + f.coef_off[0] + static_cast<size_t>(block) * 64;
  const int c = f.blk_comp[j];
  const unsigned bx = mx * f.comp_h[c] + f.blk_hh[j], by = my * f.comp_v[c] + f.blk_vv[j];
  if (bx >= static_cast<unsigned>(f.wblocks[c]) || by >= static_cast<unsigned>(f.hblocks[c])) return nullptr;
  return coefs + f.coef_off[c] + (static_cast<size_t>(by) * f.wblocks[c] + bx) * 64;
}

__device__ __forceinline__ int16_t* block_dst(const HuffFileDesc& f, int16_t* coefs, unsigned block) {
  const unsigned mcu = block / f.bpm, j = block - mcu * f.bpm;
  const int c = f.blk_comp[j];
  const unsigned my = mcu / f.mcus_x, mx = mcu - my * f.mcus_x;
  const unsigned bx = mx * f.comp_h[c] + f.blk_hh[j], by = my * f.comp_v[c] + f.blk_vv[j];
  if (f.ncomp == 1) return coefs + f.coef_off[0] + static_cast<size_t>(block) * 64;
  if (bx >= static_cast<unsigned>(f.wblocks[c]) || by >= static_cast<unsigned>(f.hblocks[c])) return nullptr;
  return coefs + f.coef_off[c] + (static_cast<size_t>(by) * f.wblocks[c] + bx) * 64;
}

// Decodes from (bit, blk, k) until the bit position reaches end_bit (or, when writing, the restart segment is complete).
// Returns the end state; *nblk = blocks completed on the way.
template <bool WRITE>
__device__ __forceinline__ unsigned long long decode_span(const SmemHuff& sm, unsigned win_bit0, unsigned bit, unsigned blk,
                                                          unsigned k, unsigned end_bit, int bpm,
                                                          unsigned long long blk_comp, unsigned* nblk, WriteCtx* wc) {
  unsigned done = 0;
  // 64-bit window on the stream: the next symbol's bits are its top 32; refilled one word at a time
  unsigned w = (bit - win_bit0) >> 5;
  const unsigned skew = (bit - win_bit0) & 31;
  unsigned long long buf = ((static_cast<unsigned long long>(__byte_perm(sm.words[w], 0, 0x0123)) << 32) |
                            __byte_perm(sm.words[w + 1], 0, 0x0123))
                           << skew;
  int avail = 64 - static_cast<int>(skew);
  w += 2;
  while (bit < end_bit) {
    const int c = static_cast<int>(blk_comp >> (8 * blk)) & 3;
    const DevHuffTable& t = sm.tab[2 * c + (k ? 1 : 0)];
    if (avail < 32) {
      buf |= static_cast<unsigned long long>(__byte_perm(sm.words[w++], 0, 0x0123)) << (32 - avail);
      avail += 32;
    }
    const unsigned bits = static_cast<unsigned>(buf >> 32);
    unsigned e = t.fast[bits >> 22];
    unsigned len, sym;
    if (e) {
      len = e >> 8;
      sym = e & 255;
    } else {
      len = 0;
      sym = 0;
#pragma unroll 1
      for (int l = 11; l <= 16; ++l) {
        const int code = static_cast<int>(bits >> (32 - l));
        if (code <= t.maxcode[l]) {
          len = l;
          sym = t.vals[code + t.valoff[l]];
          break;
        }
      }
      if (len == 0) {  // not a code: a decoder that is not synchronised yet moves on, the final pass reports damage
        if (WRITE) *wc->error = 1;
        len = 1;
      }
    }
    // One symbol, written with selects rather than branches so that the lanes of a warp stay together whatever kind of
    // symbol each of them is on: a DC symbol is an AC symbol with run 0 stored at position 0.
    const bool is_dc = k == 0;
    const unsigned r = is_dc ? 0u : sym >> 4, sz = sym & 15;
    const unsigned n = len + sz;
    const bool ends = !is_dc && sz == 0;  // end of block (run 0) or sixteen zeros (run 15)
    const unsigned kpos = k + r;          // where this coefficient goes
    if (WRITE) {
      const unsigned raw = static_cast<unsigned>((static_cast<unsigned long long>(bits << len) << sz) >> 32);
      const unsigned half = (1u << sz) >> 1;
      const int v = raw < half ? static_cast<int>(raw) - static_cast<int>(1u << sz) + 1 : static_cast<int>(raw);
      if (is_dc) {
        if (sym > 15) *wc->error = 1;
        wc->dcdiff[wc->block] = static_cast<int16_t>(v);
      } else if (!ends) {
        if (kpos > 63) *wc->error = 1;
        else if (wc->dst) wc->dst[sm.zigzag[kpos]] = static_cast<int16_t>(v);
      }
    }
    k = ends ? (r == 15 ? k + 16 : 64u) : kpos + 1;
    bit += n;
    buf <<= n;
    avail -= static_cast<int>(n);
    const bool fin = k >= 64;
    k = fin ? 0u : k;
    blk = fin ? (blk + 1 == static_cast<unsigned>(bpm) ? 0u : blk + 1) : blk;
    done += fin ? 1u : 0u;
    if (WRITE && fin) {
      ++wc->block;
      if (wc->block >= wc->seg_end) break;
      if (blk == 0) {  // next MCU
        if (++wc->mx == static_cast<unsigned>(wc->f->mcus_x)) {
          wc->mx = 0;
          ++wc->my;
        }
      }
      wc->dst = mcu_block_dst(*wc->f, wc->coefs, wc->mx, wc->my, blk, wc->block);
    }
  }
  *nblk = done;
  return pack_state(bit, blk, k);
}

__device__ __forceinline__ void load_block(SmemHuff& sm, const HuffFileDesc& f, const DevHuffTable* tables,
                                           const unsigned char* streams, unsigned first_sub) {
  const unsigned* src_t = reinterpret_cast<const unsigned*>(tables + f.table_index);
  unsigned* dst_t = reinterpret_cast<unsigned*>(sm.tab);
  for (int i = threadIdx.x; i < static_cast<int>(6 * sizeof(DevHuffTable) / 4); i += blockDim.x) dst_t[i] = src_t[i];
  if (threadIdx.x < 80) sm.zigzag[threadIdx.x] = kZigzagDev[threadIdx.x];
  // the arena carries slack behind the last stream, so the look-ahead words never leave it
  const unsigned* src = reinterpret_cast<const unsigned*>(streams + f.stream_off) + static_cast<size_t>(first_sub) * (kSubseqBytes / 4);
  const unsigned avail = (f.n_sub - first_sub) * (kSubseqBytes / 4) + 8;
  const unsigned n = min(static_cast<unsigned>(kHuffThreads * kSubseqBytes / 4 + 8), avail);
  for (unsigned i = threadIdx.x; i < n; i += blockDim.x) sm.words[i] = src[i];
  __syncthreads();
}

// Passes 1 and 2.  state[] = end state of every subsequence, start_used[] = the state its last decode started from.
__global__ void __launch_bounds__(kHuffThreads)
huff_sync_kernel(const HuffFileDesc* __restrict__ files, const HuffBlockDesc* __restrict__ blocks,
                 const DevHuffTable* __restrict__ tables, const unsigned char* __restrict__ streams,
                 const int* __restrict__ sub_seg, unsigned long long* state, unsigned long long* start_used,
                 unsigned* nblk, int first_round, int* changed) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SmemHuff& sm = *reinterpret_cast<SmemHuff*>(smem_raw);
  // per subsequence of this CUDA block: end state, the start state it was decoded from, blocks completed (bit 31 =
  // decoded in this launch); any thread may decode any subsequence (see the work list below)
  __shared__ unsigned long long s_state[kHuffThreads];
  __shared__ unsigned long long s_used[kHuffThreads];
  __shared__ unsigned s_cnt[kHuffThreads];
  __shared__ unsigned short s_list[kHuffThreads];
  __shared__ int s_n;
  constexpr unsigned kDirty = 0x80000000u;
  const HuffBlockDesc bd = blocks[blockIdx.x];
  const HuffFileDesc& f = files[bd.file];
  load_block(sm, f, tables, streams, bd.first_sub);
  const unsigned long long comp_of = *reinterpret_cast<const unsigned long long*>(f.blk_comp);  // component of block j
  const unsigned i = bd.first_sub + threadIdx.x;
  const bool active = i < f.n_sub;
  const unsigned gi = f.sub_base + i;
  const unsigned win_bit0 = bd.first_sub * kSubBits;
  const bool first = active && (i == 0 || sub_seg[gi] != sub_seg[gi - 1]);
  const unsigned long long fixed = pack_state(i * kSubBits, 0, 0);  // what a restart segment starts with
  {
    unsigned long long used = 0, mine = 0;
    unsigned cnt = 0;
    if (active) {
      if (first_round) {
        if (first) {
          used = fixed;
          mine = decode_span<false>(sm, win_bit0, i * kSubBits, 0, 0, (i + 1) * kSubBits, f.bpm, comp_of, &cnt, nullptr);
        } else {
          // the guess only has to END in the right state: decoding the last kGuessBits of the subsequence is enough
          // for the decoder to fall into step; `used` = a state no predecessor can end in, so the first sweep decodes
          // the whole subsequence from the predecessor's end state
          used = ~0ull;
          mine = decode_span<false>(sm, win_bit0, (i + 1) * kSubBits - kGuessBits, 0, 0, (i + 1) * kSubBits, f.bpm,
                                    comp_of, &cnt, nullptr);
        }
        cnt |= kDirty;
      } else {
        used = start_used[gi];
        mine = state[gi];
        cnt = nblk[gi];
      }
    }
    s_state[threadIdx.x] = mine;
    s_used[threadIdx.x] = used;
    s_cnt[threadIdx.x] = cnt;
  }
  for (int it = 0; it < kMaxInner; ++it) {
    if (threadIdx.x == 0) s_n = 0;
    __syncthreads();
    // which subsequences start from a state their predecessor no longer ends in?  They go on a work list, so that the
    // re-decoding below keeps whole warps busy even when only a few subsequences are left.
    if (active && !first) {
      // across CUDA blocks the predecessor's state comes from the previous launch (none yet in the first round)
      unsigned long long prev = s_used[threadIdx.x];
      if (threadIdx.x) prev = s_state[threadIdx.x - 1];
      else if (!first_round) prev = __ldcg(&state[gi - 1]);
      const unsigned pbit = static_cast<unsigned>(prev);
      const bool sane = pbit >= i * kSubBits && pbit < i * kSubBits + 32;  // (always; keeps reads inside the window)
      if (sane && prev != s_used[threadIdx.x]) {
        s_used[threadIdx.x] = prev;
        s_list[atomicAdd(&s_n, 1)] = static_cast<unsigned short>(threadIdx.x);
      }
    }
    __syncthreads();
    const int n = s_n;
    if (n == 0) break;
    if (static_cast<int>(threadIdx.x) < n) {
      const unsigned j = s_list[threadIdx.x];
      const unsigned long long st = s_used[j];
      unsigned cnt = 0;
      s_state[j] = decode_span<false>(sm, win_bit0, static_cast<unsigned>(st), static_cast<unsigned>(st >> 32) & 255,
                                      static_cast<unsigned>(st >> 40) & 255, (bd.first_sub + j + 1) * kSubBits, f.bpm,
                                      comp_of, &cnt, nullptr);
      s_cnt[j] = cnt | kDirty;
    }
    __syncthreads();
  }
  if (active && (s_cnt[threadIdx.x] & kDirty)) {
    state[gi] = s_state[threadIdx.x];
    start_used[gi] = s_used[threadIdx.x];
    nblk[gi] = s_cnt[threadIdx.x] & ~kDirty;
    if (!first_round) *changed = 1;
  }
}

// Pass 3a: per CUDA block, segmented exclusive scan of nblk (restart segments reset the count); block totals out.
__global__ void __launch_bounds__(kHuffThreads)
huff_scan_kernel(const HuffFileDesc* __restrict__ files, const HuffBlockDesc* __restrict__ blocks,
                 const int* __restrict__ sub_seg, const unsigned* __restrict__ nblk, unsigned* local_off,
                 unsigned* block_sum, int* block_has_start) {
  __shared__ unsigned s_val[kHuffThreads];
  __shared__ int s_flag[kHuffThreads];
  const HuffBlockDesc bd = blocks[blockIdx.x];
  const HuffFileDesc& f = files[bd.file];
  const unsigned i = bd.first_sub + threadIdx.x;
  const bool active = i < f.n_sub;
  const unsigned gi = f.sub_base + i;
  const int first = active && (i == 0 || sub_seg[gi] != sub_seg[gi - 1]);
  unsigned v = active ? nblk[gi] : 0;
  int fl = first;
  s_val[threadIdx.x] = v;
  s_flag[threadIdx.x] = fl;
  __syncthreads();
  // inclusive segmented scan (Hillis-Steele): (v, f) o (v', f') = (f' ? v' : v + v', f | f')
  for (int d = 1; d < kHuffThreads; d <<= 1) {
    unsigned pv = 0;
    int pf = 0;
    if (threadIdx.x >= static_cast<unsigned>(d)) {
      pv = s_val[threadIdx.x - d];
      pf = s_flag[threadIdx.x - d];
    }
    __syncthreads();
    if (threadIdx.x >= static_cast<unsigned>(d)) {
      if (!fl) v += pv;
      fl |= pf;
      s_val[threadIdx.x] = v;
      s_flag[threadIdx.x] = fl;
    }
    __syncthreads();
  }
  // exclusive value = blocks completed since the last segment start before this subsequence (inside this CUDA block)
  if (active) {
    const unsigned incl_prev = threadIdx.x ? s_val[threadIdx.x - 1] : 0;
    local_off[gi] = first ? 0 : incl_prev;
  }
  if (threadIdx.x == kHuffThreads - 1) {
    block_sum[blockIdx.x] = v;       // blocks since the last segment start inside this CUDA block (or since its beginning)
    block_has_start[blockIdx.x] = fl;
  }
}

// Pass 3b: carries between CUDA blocks: carry[b] = blocks completed since the last restart-segment start before CUDA
// block b (0 at the first block of a file).  One thread block, chunked segmented scan.
__global__ void __launch_bounds__(1024)
huff_carry_kernel(const HuffBlockDesc* __restrict__ blocks, int n_blocks, const unsigned* __restrict__ block_sum,
                  const int* __restrict__ block_has_start, unsigned* carry) {
  __shared__ unsigned s_val[1024];
  __shared__ int s_flag[1024];
  const int chunk = (n_blocks + 1023) / 1024;
  const int lo = min(static_cast<int>(threadIdx.x) * chunk, n_blocks), hi = min(lo + chunk, n_blocks);
  // a CUDA block restarts the count when it holds a segment start or is the first of its file
  auto restarts = [&](int b) { return block_has_start[b] || b == 0 || blocks[b].file != blocks[b - 1].file; };
  unsigned run = 0;
  int flag = 0;
  for (int b = lo; b < hi; ++b) {
    if (restarts(b)) {
      run = 0;
      flag = 1;
    }
    run += block_sum[b];
  }
  s_val[threadIdx.x] = run;
  s_flag[threadIdx.x] = flag;
  __syncthreads();
  unsigned v = run;
  int fl = flag;
  for (int d = 1; d < 1024; d <<= 1) {
    unsigned pv = 0;
    int pf = 0;
    if (threadIdx.x >= static_cast<unsigned>(d)) {
      pv = s_val[threadIdx.x - d];
      pf = s_flag[threadIdx.x - d];
    }
    __syncthreads();
    if (threadIdx.x >= static_cast<unsigned>(d)) {
      if (!fl) v += pv;
      fl |= pf;
      s_val[threadIdx.x] = v;
      s_flag[threadIdx.x] = fl;
    }
    __syncthreads();
  }
  run = threadIdx.x ? s_val[threadIdx.x - 1] : 0;  // count entering this thread's chunk
  for (int b = lo; b < hi; ++b) {
    const bool new_file = b == 0 || blocks[b].file != blocks[b - 1].file;
    if (new_file) run = 0;
    carry[b] = run;  // for the subsequences of block b that precede its first segment start
    run = (block_has_start[b] ? 0 : run) + block_sum[b];
  }
}

// Pass 4: decode from the synchronised states and write.
__global__ void __launch_bounds__(kHuffThreads)
huff_write_kernel(const HuffFileDesc* __restrict__ files, const HuffBlockDesc* __restrict__ blocks,
                  const DevHuffTable* __restrict__ tables, const unsigned char* __restrict__ streams,
                  const int* __restrict__ sub_seg, const unsigned long long* __restrict__ start_used,
                  const unsigned* __restrict__ local_off, const unsigned* __restrict__ carry, int16_t* coefs,
                  int16_t* dcdiff, int* file_error) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SmemHuff& sm = *reinterpret_cast<SmemHuff*>(smem_raw);
  __shared__ int s_seen_start;
  const HuffBlockDesc bd = blocks[blockIdx.x];
  const HuffFileDesc& f = files[bd.file];
  if (threadIdx.x == 0) s_seen_start = kHuffThreads;
  load_block(sm, f, tables, streams, bd.first_sub);
  const unsigned long long comp_of = *reinterpret_cast<const unsigned long long*>(f.blk_comp);
  const unsigned i = bd.first_sub + threadIdx.x;
  const bool active = i < f.n_sub;
  const unsigned gi = f.sub_base + i;
  const bool first = active && (i == 0 || sub_seg[gi] != sub_seg[gi - 1]);
  if (first) atomicMin(&s_seen_start, static_cast<int>(threadIdx.x));
  __syncthreads();
  if (!active) return;
  const unsigned seg = static_cast<unsigned>(sub_seg[gi]);
  const unsigned seg_first_block = seg * f.seg_blocks;
  const unsigned seg_end = min(seg_first_block + f.seg_blocks, f.total_blocks);
  // blocks completed in this restart segment before this subsequence
  unsigned before = local_off[gi];
  if (static_cast<int>(threadIdx.x) < s_seen_start) before += carry[blockIdx.x];
  const unsigned long long st = start_used[gi];
  WriteCtx wc;
  wc.f = &f;
  wc.coefs = coefs;
  wc.dcdiff = dcdiff + f.dc_off;
  wc.block = seg_first_block + before;
  wc.seg_end = seg_end;
  wc.error = file_error + bd.file;
  const bool last_of_seg = i + 1 == f.n_sub || sub_seg[gi + 1] != sub_seg[gi];
  if (wc.block >= seg_end) {
    // nothing left for this subsequence (padding behind the last block of the segment) - unless blocks were lost
    return;
  }
  {
    const unsigned mcu = wc.block / f.bpm;
    wc.my = mcu / f.mcus_x;
    wc.mx = mcu - wc.my * f.mcus_x;
  }
  wc.dst = block_dst(f, coefs, wc.block);
  unsigned cnt = 0;
  decode_span<true>(sm, bd.first_sub * kSubBits, static_cast<unsigned>(st), static_cast<unsigned>(st >> 32) & 255,
                    static_cast<unsigned>(st >> 40) & 255, (i + 1) * kSubBits, f.bpm, comp_of, &cnt, &wc);
  if (last_of_seg && wc.block < seg_end) *wc.error = 1;  // the data ended before the segment's last block
}

// DC prediction: per (file, component) an inclusive sum of the differences in scan order, reset at restart segments.
// The sequence of a component is cut into kDcParts parts, one CUDA block each: a first kernel reduces every part to
// (sum since the last reset, reset seen), the second scans inside the parts with the carry of the parts before.
constexpr int kDcParts = 8;
constexpr int kDcThreads = 256;

struct DcRange {
  const int16_t* dd;
  int j0, per;
  unsigned seg_len, lo, hi;  // this thread's elements [lo, hi) of the component's sequence
  bool valid;
};
__device__ __forceinline__ DcRange dc_range(const HuffFileDesc& f, const int16_t* dcdiff, int c, int part) {
  DcRange r{};
  r.valid = c < f.ncomp;
  if (!r.valid) return r;
  // blocks of component c inside an MCU are consecutive: [j0, j0 + per)
  r.j0 = 0;
  for (int cc = 0; cc < c; ++cc) r.j0 += f.comp_h[cc] * f.comp_v[cc];
  r.per = f.ncomp == 1 ? 1 : f.comp_h[c] * f.comp_v[c];
  const unsigned mcus = f.total_blocks / f.bpm;
  const unsigned n = mcus * r.per;                 // blocks of this component in scan order
  r.seg_len = (f.seg_blocks / f.bpm) * r.per;      // ... per restart segment
  r.dd = dcdiff + f.dc_off;
  const unsigned plen = (n + kDcParts - 1) / kDcParts;
  const unsigned plo = min(part * plen, n), phi = min(plo + plen, n);
  const unsigned chunk = (phi - plo + kDcThreads - 1) / kDcThreads;
  r.lo = min(plo + threadIdx.x * chunk, phi);
  r.hi = min(r.lo + chunk, phi);
  return r;
}
// (sum since the last reset, reset seen) of this thread's elements; element t = block (t / per) * bpm + j0 + t % per
__device__ __forceinline__ void dc_chunk_sum(const HuffFileDesc& f, const DcRange& r, int* sum_out, int* flag_out) {
  unsigned mcu = r.lo / r.per, q = r.lo - mcu * r.per, left = r.lo % r.seg_len ? r.seg_len - r.lo % r.seg_len : 0;
  int sum = 0, flag = 0;
  for (unsigned t = r.lo; t < r.hi; ++t) {
    if (left == 0) {
      sum = 0;
      flag = 1;
      left = r.seg_len;
    }
    --left;
    sum += r.dd[mcu * f.bpm + r.j0 + q];
    if (++q == static_cast<unsigned>(r.per)) {
      q = 0;
      ++mcu;
    }
  }
  *sum_out = sum;
  *flag_out = flag;
}
// inclusive segmented scan over the block's threads; returns this thread's inclusive (sum, flag)
__device__ __forceinline__ void dc_block_scan(int* s_sum, int* s_flag, int* v_io, int* fl_io) {
  int v = *v_io, fl = *fl_io;
  s_sum[threadIdx.x] = v;
  s_flag[threadIdx.x] = fl;
  __syncthreads();
  for (int d = 1; d < kDcThreads; d <<= 1) {
    int pv = 0, pf = 0;
    if (threadIdx.x >= static_cast<unsigned>(d)) {
      pv = s_sum[threadIdx.x - d];
      pf = s_flag[threadIdx.x - d];
    }
    __syncthreads();
    if (threadIdx.x >= static_cast<unsigned>(d)) {
      if (!fl) v += pv;
      fl |= pf;
      s_sum[threadIdx.x] = v;
      s_flag[threadIdx.x] = fl;
    }
    __syncthreads();
  }
  *v_io = v;
  *fl_io = fl;
}

// grid: n_files * 3 * kDcParts; part_sum / part_flag: one entry per CUDA block
__global__ void __launch_bounds__(kDcThreads)
huff_dc_part_kernel(const HuffFileDesc* __restrict__ files, const int16_t* __restrict__ dcdiff, int* part_sum,
                    int* part_flag) {
  __shared__ int s_sum[kDcThreads];
  __shared__ int s_flag[kDcThreads];
  const int part = blockIdx.x % kDcParts, fc = blockIdx.x / kDcParts;
  const HuffFileDesc& f = files[fc / 3];
  const DcRange r = dc_range(f, dcdiff, fc % 3, part);
  if (!r.valid) return;
  int v, fl;
  dc_chunk_sum(f, r, &v, &fl);
  dc_block_scan(s_sum, s_flag, &v, &fl);
  if (threadIdx.x == kDcThreads - 1) {
    part_sum[blockIdx.x] = v;
    part_flag[blockIdx.x] = fl;
  }
}

__global__ void __launch_bounds__(kDcThreads)
huff_dc_kernel(const HuffFileDesc* __restrict__ files, const int16_t* __restrict__ dcdiff, const int* __restrict__ part_sum,
               const int* __restrict__ part_flag, int16_t* coefs) {
  __shared__ int s_sum[kDcThreads];
  __shared__ int s_flag[kDcThreads];
  const int part = blockIdx.x % kDcParts, fc = blockIdx.x / kDcParts;
  const HuffFileDesc& f = files[fc / 3];
  const DcRange r = dc_range(f, dcdiff, fc % 3, part);
  if (!r.valid) return;
  // predictor entering this part: the parts before it, combined in order
  int enter = 0;
  for (int p = 0; p < part; ++p) {
    const int b = blockIdx.x - part + p;
    enter = part_flag[b] ? part_sum[b] : enter + part_sum[b];
  }
  int v, fl;
  dc_chunk_sum(f, r, &v, &fl);
  if (threadIdx.x == 0 && !fl) v += enter;
  dc_block_scan(s_sum, s_flag, &v, &fl);
  int pred = threadIdx.x ? s_sum[threadIdx.x - 1] : enter;  // predictor entering this thread's chunk
  unsigned mcu = r.lo / r.per, q = r.lo - mcu * r.per, left = r.lo % r.seg_len ? r.seg_len - r.lo % r.seg_len : 0;
  unsigned my = mcu / f.mcus_x, mx = mcu - my * f.mcus_x;
  for (unsigned t = r.lo; t < r.hi; ++t) {
    if (left == 0) {
      pred = 0;
      left = r.seg_len;
    }
    --left;
    const unsigned b = mcu * f.bpm + r.j0 + q;
    pred += r.dd[b];
    int16_t* dst = mcu_block_dst(f, coefs, mx, my, r.j0 + q, b);
    if (dst) dst[0] = static_cast<int16_t>(pred);
    if (++q == static_cast<unsigned>(r.per)) {
      q = 0;
      ++mcu;
      if (++mx == static_cast<unsigned>(f.mcus_x)) {
        mx = 0;
        ++my;
      }
    }
  }
}

}  // namespace

cudaError_t HuffDecode(const HuffBatch& b, cudaStream_t st) {
  if (b.n_blocks <= 0) return cudaSuccess;
  static bool attr_set[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 64 && !attr_set[dev]) {
    cudaFuncSetAttribute(huff_sync_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sizeof(SmemHuff)));
    cudaFuncSetAttribute(huff_write_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sizeof(SmemHuff)));
    attr_set[dev] = true;
  }
  const size_t smem = sizeof(SmemHuff);
  huff_sync_kernel<<<b.n_blocks, kHuffThreads, smem, st>>>(b.files, b.blocks, b.tables, b.streams, b.sub_seg, b.state,
                                                          b.start_used, b.nblk, 1, b.changed);
  // re-synchronisation rounds: three launches in a row, then one at a time until a launch changed nothing
  int rounds = 0;
  for (;;) {
    const int burst = rounds == 0 ? 3 : 1;
    cudaError_t e = cudaMemsetAsync(b.changed, 0, sizeof(int), st);
    if (e != cudaSuccess) return e;
    for (int r = 0; r < burst; ++r) {
      if (r == burst - 1 && burst > 1) {
        e = cudaMemsetAsync(b.changed, 0, sizeof(int), st);
        if (e != cudaSuccess) return e;
      }
      huff_sync_kernel<<<b.n_blocks, kHuffThreads, smem, st>>>(b.files, b.blocks, b.tables, b.streams, b.sub_seg,
                                                              b.state, b.start_used, b.nblk, 0, b.changed);
    }
    rounds += burst;
    e = cudaMemcpyAsync(b.h_changed, b.changed, sizeof(int), cudaMemcpyDeviceToHost, st);
    if (e != cudaSuccess) return e;
    e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return e;
    if (*b.h_changed == 0) break;
    if (rounds >= kMaxRounds) {
      // a stream whose decoders do not re-synchronise (one CUDA block settles per launch at worst): not worth chasing
      // on the device - the caller decodes this batch's files with the sequential host decoder
      if (b.rounds_out) *b.rounds_out = -1;
      return cudaSuccess;
    }
  }
  huff_scan_kernel<<<b.n_blocks, kHuffThreads, 0, st>>>(b.files, b.blocks, b.sub_seg, b.nblk, b.local_off, b.block_sum,
                                                        b.block_has_start);
  huff_carry_kernel<<<1, 1024, 0, st>>>(b.blocks, b.n_blocks, b.block_sum, b.block_has_start, b.carry);
  huff_write_kernel<<<b.n_blocks, kHuffThreads, smem, st>>>(b.files, b.blocks, b.tables, b.streams, b.sub_seg,
                                                           b.start_used, b.local_off, b.carry, b.coefs, b.dcdiff,
                                                           b.file_error);
  huff_dc_part_kernel<<<b.n_files * 3 * kDcParts, kDcThreads, 0, st>>>(b.files, b.dcdiff, b.dc_part, b.dc_part + b.n_files * 3 * kDcParts);
  huff_dc_kernel<<<b.n_files * 3 * kDcParts, kDcThreads, 0, st>>>(b.files, b.dcdiff, b.dc_part, b.dc_part + b.n_files * 3 * kDcParts, b.coefs);
  if (b.rounds_out) *b.rounds_out = rounds;
  return cudaGetLastError();
}

}  // namespace rn
