// 16-bit tensor-core path for sm_100a: tcgen05.mma (kind::f16, fp32 accumulate in
// TMEM) implicit-GEMM 3x3 VALID convolution with a fused bias + ReLU6 + avg-pool
// epilogue, fed by 1-D TMA bulk copies (cp.async.bulk + mbarrier complete_tx).
//
// Reference semantics: one depth step of conv_block (reference network.py:184-194)
// in folded form  p' = pool(relu6(conv_W(p) + b)).
//
// Design ("row-stationary tap stacking"):
//   * A CTA walks a 128-pixel-wide column strip of one image top to bottom.
//   * M tile   = 128 consecutive pixels of ONE input row (UMMA M = 128, lane = pixel).
//   * A operand = that input row in shared memory, channel-chunk planes
//                 [cb][pixel][8 ch] = the UMMA no-swizzle K-major core-matrix order;
//                 the dx tap shift is a +16 B*dx descriptor start offset.
//   * B operand = the weights of all three dy taps stacked along N:
//                 [W(dy=2) | W(dy=1) | W(dy=0)]  (N = 3*Cout).
//   * D         = the accumulators of conv rows r-2, r-1, r, adjacent TMEM column
//                 blocks of a ring of R = 512/Cout row slots.
//   Each input row is therefore read from shared memory 3*(Cin/16) times (not 9x)
//   and from L2/HBM exactly once per strip; N grows from Cout to 3*Cout which is
//   what makes Cout = 32/64 layers feed the tensor pipe.
//   * Epilogue warps read a finished conv row from TMEM (lane = pixel, column =
//     channel), add bias, clip, keep the vertical pooling window in registers, do
//     the horizontal window with warp shuffles, and write 16-byte channel chunks.
#include <cstdio>
#include <cstring>
#include <vector>

#include "kernels.h"
#include "tc_common.cuh"

namespace rn {

namespace {

struct TcParams {
  const uint8_t* in;
  uint8_t* out;
  const uint8_t* w;  // packed, per part
  const float* bias;
  int N;
  int in_side, conv_side, out_side;
  int cb_out_total;   // channel chunks of the full output tensor
  int n_strips;
  int strip_step_in;  // input/conv column step between strips
  int strip_step_out; // output column step between strips
  int win_step_in;    // POOL != 0: input-pixel step between the four 32-lane windows of a tile
  int win_step_out;   // ... and the number of output columns each window produces
  int total_rows;     // work line: image groups x out_side pooled rows (tc_common.cuh: rn_gang_rows)
  int min_piece;      // ... and the snapping distance of the CTA ranges
  int w_bytes;        // packed weight bytes per part
  // fused residual join (JOIN kernels): out = A*h + B*resize_bilinear_legacy(src)[y][x] + C
  const uint8_t* res_src;  // chunked tensor [n][y][c/8][x][8], side res_side, same channel count as the output
  int res_side;
  float res_scale;         // res_side / out_side (float32, as TF computes it)
  const float* join_abc;   // device [3][cout]: A, B, C (A, B already divided by the stored-activation scales)
  int dbg;                 // timing experiments only (RN_TC_DBG): 1 = no stores, 2 = no pooling math, 4 = no TMEM re-init
                           // (honoured only in builds with -DRN_TC_TIMING_EXPERIMENTS)
};


struct Item {
  int n0, strip, x_in0, x_out0, po0, npo, c0, nconv;
};

// the piece of the CTA's row range [.., hi) that starts at line position `cur`
template <int POOL, int SEG>
__device__ __forceinline__ Item decode_piece(const TcParams& p, int cur, int hi) {
  Item it;
  const int ig = cur / p.out_side;
  const int strip = blockIdx.x % p.n_strips;
  it.n0 = ig * SEG;
  it.strip = strip;
  it.x_in0 = strip * p.strip_step_in;
  it.x_out0 = strip * p.strip_step_out;
  it.po0 = cur - ig * p.out_side;
  it.npo = min(p.out_side - it.po0, hi - cur);
  if (POOL == 0) {
    it.c0 = it.po0;
    it.nconv = it.npo;
  } else if (POOL == 42) {
    it.c0 = it.po0 * 2;
    it.nconv = (it.npo - 1) * 2 + 4;
  } else {
    it.c0 = it.po0;
    it.nconv = it.npo + (POOL == 41 ? 3 : 2);
  }
  // the epilogue consumes two conv rows per iteration: an odd count gets one extra (never stored) row, whose
  // input rows may lie past the image (next image or the slack behind the tensor) - harmless garbage
  it.nconv = (it.nconv + 1) & ~1;
  return it;
}


// ---------------------------------------------------------------------------
// CB   : input channel chunks (Cin/8)         COUT : output channels of this pass
// POOL : 0 / 31 / 41 / 42                     SEG  : images side by side in one 128-pixel tile
// ---------------------------------------------------------------------------
// CREAL: channels actually produced (<= COUT; conv0 pads 8 -> 16 to satisfy UMMA N % 16 == 0)
// horizontal leg of the residual resize for one source row: left/right taps of CG channels (16-byte chunks one plane
// apart), out[i] = l + (r - l) * tx on packed 16-bit pairs
template <typename HH, int CG>
__device__ __forceinline__ void join_hlerp(const uint8_t* row, uint32_t dx_bytes, uint32_t plane_bytes, uint32_t tx2,
                                           uint32_t* out) {
#pragma unroll
  for (int cb = 0; cb < CG / 8; ++cb) {
    const uint4 l = *reinterpret_cast<const uint4*>(row + cb * plane_bytes);
    const uint4 r = *reinterpret_cast<const uint4*>(row + cb * plane_bytes + dx_bytes);
    out[4 * cb + 0] = HH::fma(HH::sub(r.x, l.x), tx2, l.x);
    out[4 * cb + 1] = HH::fma(HH::sub(r.y, l.y), tx2, l.y);
    out[4 * cb + 2] = HH::fma(HH::sub(r.z, l.z), tx2, l.z);
    out[4 * cb + 3] = HH::fma(HH::sub(r.w, l.w), tx2, l.w);
  }
}

#ifdef RN_TC_TIMING_EXPERIMENTS
#define RN_TC_DBG_BIT(p, bit) (((p).dbg & (bit)) != 0)
#else
#define RN_TC_DBG_BIT(p, bit) false
#endif

// SPLIT: the fp32-class path (RN_PREC_FP32_TC).  Activations travel as hi + lo 16-bit pairs (planes [hi | lo], CB counts
// the physical planes), every input row takes the three products xh*Wh + xl*Wh + xh*Wl into the same fp32 accumulators
// (TcCfg), and the epilogue keeps everything in fp32 - pooling windows, residual join - until it splits the result into
// hi = round16(v), lo = round16(v - hi) for the store.
template <int CB, int COUT, int POOL, int SEG, int AMODE, bool BF16, int CREAL = COUT, bool JOIN = false, bool SPLIT = false>
__global__ void __launch_bounds__(tc_threads(CREAL), 1) conv_tc_kernel(const TcParams p, const __grid_constant__ CUtensorMap tmap) {
  using Cfg = TcCfg<CB, COUT, AMODE, POOL != 0, SPLIT && AMODE == 0>;
  static_assert(!SPLIT || (CREAL % 8 == 0 && (CREAL / tc_groups(CREAL)) % 8 == 0), "split stores write whole 16-byte chunks");
  static_assert(POOL == 0 || SEG == 1 || POOL == 42, "two-image windowed tiles exist for 4x4/2 pooling only");
  using HH = H2<BF16>;
  constexpr int R = Cfg::kSlots;
  constexpr int LOGR = Cfg::kLogSlots;
  constexpr int NST = Cfg::kStages;
  constexpr int SEGW = kTileM / SEG;
  constexpr int NG = tc_groups(CREAL);     // epilogue channel groups (4 warps each)
  constexpr int kThreadsTc = tc_threads(CREAL);
  constexpr int CG = CREAL / NG;           // channels per epilogue group
  constexpr int NP = CG / 2;     // half2 pairs per thread
  // POOL != 0: the 128 lanes are four 32-pixel windows that overlap in the image (each window carries its own
  // pooling halo), so no epilogue warp ever needs a neighbour quadrant's columns.  POOL == 0: one contiguous run.
  constexpr bool kWindows = POOL != 0;
  // L2 prefetch distance of the TMA producer in row pairs (0 = off).  Measured per layer: it helps the tensor-bound
  // 32->64 layer (0.333 -> 0.315 ms per 256 images) and costs 2-7 % on the layers that already run near the HBM or
  // issue limits, so only that instantiation uses it.
  constexpr int kPrefetchPairs = (CB == 4 && COUT == 64 && !JOIN) ? 6 : 0;
  constexpr uint32_t kStageTx = 2 * (kWindows ? CB * 4 * 32 * 16 : CB * (SEG == 1 ? kLoadPx * 16 : 2 * SEGW * 16));
  constexpr uint32_t kDescHi = (128u >> 4) | (1u << 14);  // SBO = 128 B, descriptor version 1, SWIZZLE_NONE

  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* s_w = smem;
  uint8_t* s_stage = smem + Cfg::kWBytes;
  float* s_bias = reinterpret_cast<float*>(s_stage + NST * Cfg::kStageBytes);
  float* s_abc = s_bias + COUT;  // [3][COUT] join coefficients (JOIN kernels); the B row is re-used for ...
  uint32_t* s_bh = reinterpret_cast<uint32_t*>(s_abc + COUT);  // ... B as packed 16-bit pairs [COUT / 2]
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_bias + 4 * COUT);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_bar + 2 * NST + 1 + R);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar0 = smem_u32(s_bar);
  const uint32_t bar_full0 = bar0, bar_empty0 = bar0 + 8u * NST, bar_w = bar0 + 8u * (2 * NST);
  // accumulator barriers work on PAIRS of conv rows (the epilogue consumes two rows per iteration)
  constexpr int RP = R / 2;
  const uint32_t bar_accf0 = bar0 + 8u * (2 * NST + 1), bar_acce0 = bar_accf0 + 8u * RP;

  const int part = blockIdx.y;
  const uint8_t* w_gmem = p.w + static_cast<size_t>(part) * p.w_bytes;
  const float* bias_g = p.bias + part * COUT;

  if (threadIdx.x == 0) {
    for (int s = 0; s < NST; ++s) {
      mbar_init(bar_full0 + 8u * s, 1);
      mbar_init(bar_empty0 + 8u * s, 1);
    }
    mbar_init(bar_w, 1);
    for (int s = 0; s < RP; ++s) {
      mbar_init(bar_accf0 + 8u * s, 1);
      mbar_init(bar_acce0 + 8u * s, 4 * NG);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)),
                 "r"(Cfg::kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = threadIdx.x; i < COUT; i += kThreadsTc) s_bias[i] = bias_g[i];
  if (JOIN) {
    for (int i = threadIdx.x; i < 3 * COUT; i += kThreadsTc) {
      if (!SPLIT && i / COUT == 1) continue;  // (split layers keep B in fp32)
      s_abc[i] = p.join_abc[(i / COUT) * (COUT * gridDim.y) + part * COUT + (i % COUT)];
    }
    if (!SPLIT)
      for (int i = threadIdx.x; i < COUT / 2; i += kThreadsTc) {
        const float* bsrc = p.join_abc + COUT * gridDim.y + part * COUT + 2 * i;
        s_bh[i] = HH::pack(bsrc[0], bsrc[1]);  // exact: engine.cu made B a 16-bit value
      }
  }
  if (POOL != 0) {
    // The 128-byte pad behind the last plane of every stage is never written by the TMA box but is read (with
    // zero weights, Cin = 8 layers) by the tap shift of lane 125: it must hold finite values, so zero it once.
    for (int i = threadIdx.x; i < NST * 32; i += kThreadsTc)
      reinterpret_cast<uint32_t*>(s_stage + (i / 32) * Cfg::kStageBytes + 2 * Cfg::kRowBytes)[i % 32] = 0u;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;
  pdl_trigger();  // all CTAs of this persistent grid are resident: the next kernel may take SMs as they free up

  const size_t in_row_bytes = static_cast<size_t>(CB) * p.in_side * 16;
  const size_t in_img_bytes = in_row_bytes * p.in_side;
  int row_lo, row_hi;  // this CTA's share of the work line
  rn_gang_rows(blockIdx.x / p.n_strips, gridDim.x / p.n_strips, p.total_rows, p.out_side, p.min_piece, &row_lo, &row_hi);

  if (warp == 0) {
    // =========================== TMA producer ===========================
    // All 32 lanes issue bulk copies (one (plane, window) pair per lane and round) so that the 4*CB small
    // copies of a windowed row go out in parallel; lane 0 arms the transaction barrier first.
    if (lane == 0) {
      mbar_arrive_expect_tx(bar_w, Cfg::kWBytes);
      for (int off = 0; off < Cfg::kWBytes; off += 16384) {
        int sz = min(16384, Cfg::kWBytes - off);
        tma_bulk_g2s(smem_u32(s_w + off), w_gmem + off, sz, bar_w);
      }
    }
    pdl_wait();  // the weights are constants; the activations below are the previous kernel's output
    uint32_t st = 0, ph = 1;  // waiting parity 1 on a fresh "empty" barrier passes immediately
    uint32_t Gp = 0;          // global conv-row counter (same sequence as the MMA issuer's G)
    const uint32_t stage0 = smem_u32(s_stage);
    if constexpr (kWindows) {
      // one tiled TMA box per input row: [CB planes][4 windows][32 pixels][8 channels]; the window dimension has
      // a global stride of win_step pixels (< 32), i.e. the windows overlap in memory and every window arrives
      // with its own pooling halo
      if (lane == 0) {
        for (int cur = row_lo; cur < row_hi;) {
          const Item it = decode_piece<POOL, SEG>(p, cur, row_hi);
          cur += it.npo;
          const int nin = it.nconv + 2;  // even
          // SEG == 1: rows of all images form one dimension; SEG == 2: (image, row) are separate dimensions
          int row = SEG == 2 ? it.c0 : it.n0 * p.in_side + it.c0;
          for (int r = 0; r < nin; r += 2, row += 2) {
            if (r < it.nconv) {  // see the MMA issuer: the accumulators this input pair starts must be free
              const uint32_t gy = Gp + r;
              mbar_wait(bar_acce0 + 8u * ((gy >> 1) & (RP - 1)), (gy >> LOGR) & 1);
            }
            mbar_wait(bar_empty0 + 8u * st, ph);
            const uint32_t full = bar_full0 + 8u * st;
            mbar_arrive_expect_tx(full, kStageTx);
            if constexpr (SEG == 2) {  // two windows of images n0, n0 + 1 (past the batch / the image: zero fill)
              tma_tensor5_g2s(stage0 + st * Cfg::kStageBytes, &tmap, 0, 0, it.n0, 0, row, full);
            } else {
              tma_tensor4_g2s(stage0 + st * Cfg::kStageBytes, &tmap, 0, 4 * it.strip, 0, row, full);
              // the ring lets this thread run 7 row pairs ahead of the epilogue; the rows after that are pulled into L2
              // now, so that the load itself rarely waits for HBM
              if (kPrefetchPairs > 0 && r + 2 * kPrefetchPairs < nin)
                tma_prefetch4(&tmap, 0, 4 * it.strip, 0, row + 2 * kPrefetchPairs);
            }
            if (++st == NST) {
              st = 0;
              ph ^= 1;
            }
          }
          Gp += it.nconv;
        }
      }
    } else {
      constexpr int kCopies = 2 * SEG * CB;  // two input rows per stage
      for (int cur = row_lo; cur < row_hi;) {
        const Item it = decode_piece<POOL, SEG>(p, cur, row_hi);
        cur += it.npo;
        const int nin = it.nconv + 2;
        const uint8_t* src0 = p.in + it.n0 * in_img_bytes + it.c0 * in_row_bytes + static_cast<size_t>(it.x_in0) * 16;
        const uint8_t* src1 = p.in + min(it.n0 + 1, p.N - 1) * in_img_bytes + it.c0 * in_row_bytes;
        // per-lane copy descriptors (fixed for the whole item): source, destination offset, size
        const uint8_t* lsrc[(kCopies + 31) / 32];
        uint32_t ldst[(kCopies + 31) / 32];
#pragma unroll
        for (int k = 0; k < (kCopies + 31) / 32; ++k) {
          const int idx = min(lane + 32 * k, kCopies - 1);
          const int rr = idx / (SEG * CB), c = (idx / SEG) % CB, sg = idx % SEG;
          lsrc[k] = (sg ? src1 : src0) + rr * in_row_bytes + static_cast<size_t>(c) * p.in_side * 16;
          ldst[k] = rr * Cfg::kRowBytes + c * Cfg::kPlaneBytesT + sg * SEGW * 16;
        }
        for (int r = 0; r < nin; r += 2) {
          if (r < it.nconv) {
            const uint32_t gy = Gp + r;
            mbar_wait(bar_acce0 + 8u * ((gy >> 1) & (RP - 1)), (gy >> LOGR) & 1);
          }
          mbar_wait(bar_empty0 + 8u * st, ph);
          const uint32_t full = bar_full0 + 8u * st;
          if (lane == 0) mbar_arrive_expect_tx(full, kStageTx);
          __syncwarp();
          const uint32_t dst = stage0 + st * Cfg::kStageBytes;
#pragma unroll
          for (int k = 0; k < (kCopies + 31) / 32; ++k) {
            if (lane + 32 * k < kCopies) tma_bulk_g2s(dst + ldst[k], lsrc[k], SEG == 1 ? kLoadPx * 16 : SEGW * 16, full);
            lsrc[k] += 2 * in_row_bytes;
          }
          if (++st == NST) {
            st = 0;
            ph ^= 1;
          }
        }
        Gp += it.nconv;
      }
    }
  } else if (warp == 1) {
    // ============================ MMA issuer ============================
    // Everything here is warp-uniform; only the tcgen05 instructions themselves are issued by lane 0.
    if (elect_one()) {
    mbar_wait(bar_w, 0);
    const uint32_t a_lo0 = (smem_u32(s_stage) >> 4) | (Cfg::kALbo16 << 16);
    const uint32_t b_lo0 = (smem_u32(s_w) >> 4) | (Cfg::kBLbo16 << 16);
    const uint32_t idesc0 = make_idesc(0, BF16 ? 1 : 0);
    uint32_t st = 0, ph = 0, G = 0;
    for (int cur = row_lo; cur < row_hi;) {
      const Item it = decode_piece<POOL, SEG>(p, cur, row_hi);
      cur += it.npo;
      const int nin = it.nconv + 2;
      for (int r0 = 0; r0 < nin; r0 += 2) {
        // The accumulators of conv rows r0, r0+1 (started by this pair) are free: the producer waited for their
        // "drained + bias re-initialised" barrier BEFORE issuing the TMA that completes full[st] (one wait less on
        // this warp's critical path; mbarrier arrive/wait are release/acquire, so the ordering is transitive).
        mbar_wait(bar_full0 + 8u * st, ph);
        tc_fence_after();
        // Interior pair (both rows carry all three dy taps) whose six accumulator slots do not wrap around the ring:
        // everything but the stage and slot base is a compile-time constant - 12 bare MMAs.
        const uint32_t sb0 = (G + r0 - 2) & (R - 1);
        if (r0 >= 2 && r0 + 2 <= it.nconv && sb0 + 4 <= static_cast<uint32_t>(R)) {
          const uint32_t a0 = a_lo0 + st * (Cfg::kStageBytes >> 4);
          const uint32_t d0 = tmem_base + sb0 * COUT;
          constexpr uint32_t kIdescFull = static_cast<uint32_t>((3 * COUT) >> 3) << 17;
#pragma unroll
          for (int ks = 0; ks < Cfg::kKSteps; ++ks)
            tc_mma_acc1(d0, a0 + Cfg::a_off16(ks), b_lo0 + Cfg::b_off16(ks), kDescHi, idesc0 | kIdescFull);
#pragma unroll
          for (int ks = 0; ks < Cfg::kKSteps; ++ks)
            tc_mma_acc1(d0 + COUT, a0 + (Cfg::kRowBytes >> 4) + Cfg::a_off16(ks), b_lo0 + Cfg::b_off16(ks), kDescHi,
                        idesc0 | kIdescFull);
        } else {
#pragma unroll
          for (int sub = 0; sub < 2; ++sub) {
            const int r = r0 + sub;
            const int jlo = max(0, 2 - r);  // j = 2 - dy ; conv row y = r - 2 + j
            const int jhi = min(2, it.nconv + 1 - r);
            const uint32_t sb = (G + r - 2 + jlo) & (R - 1);
            const int nj = jhi - jlo + 1;
            const int len1 = min(nj, R - static_cast<int>(sb));  // slots before the ring wraps
            const uint32_t a_lo = a_lo0 + st * (Cfg::kStageBytes >> 4) + sub * (Cfg::kRowBytes >> 4);
            {
              const uint32_t d = tmem_base + sb * COUT;
              const uint32_t idesc = idesc0 | (static_cast<uint32_t>((len1 * COUT) >> 3) << 17);
              const uint32_t b_lo = b_lo0 + jlo * COUT;
#pragma unroll
              for (int ks = 0; ks < Cfg::kKSteps; ++ks)
                tc_mma_acc1(d, a_lo + Cfg::a_off16(ks), b_lo + Cfg::b_off16(ks), kDescHi, idesc);
            }
            if (len1 < nj) {  // ring wrap: the remaining conv rows start again at slot 0
              const uint32_t idesc = idesc0 | (static_cast<uint32_t>(((nj - len1) * COUT) >> 3) << 17);
              const uint32_t b_lo = b_lo0 + (jlo + len1) * COUT;
#pragma unroll
              for (int ks = 0; ks < Cfg::kKSteps; ++ks)
                tc_mma_acc1(tmem_base, a_lo + Cfg::a_off16(ks), b_lo + Cfg::b_off16(ks), kDescHi, idesc);
            }
          }
        }
        tc_commit(bar_empty0 + 8u * st);
        // conv rows r0-2, r0-1 received their last (dy = 2) contribution from this pair of input rows
        if (r0 >= 2) tc_commit(bar_accf0 + 8u * (((G + r0 - 2) >> 1) & (RP - 1)));
        if (++st == NST) {
          st = 0;
          ph ^= 1;
        }
      }
      G += it.nconv;
    }
    }
    __syncwarp();
  } else if (warp < 2 + 4 * NG) {
    // ============================= epilogue =============================
    const int ew = warp - 2;
    const int grp = ew >> 2;                      // channel group: channels [grp*CG, grp*CG + CG)
    const int quad = warp & 3;                    // TMEM lane quadrant this warp may access
    const int pix = quad * 32 + lane;             // pixel (UMMA row) owned by this thread
    const int seg = pix / SEGW, xs = pix % SEGW;
    const uint32_t t_base = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + grp * CG;
    const int out_planes = (SPLIT ? 2 : 1) * p.cb_out_total;  // split tensors: hi planes, then lo planes
    const size_t out_row_bytes = static_cast<size_t>(out_planes) * p.out_side * 16;
    const size_t out_img_bytes = out_row_bytes * p.out_side;
    const size_t out_plane_bytes = static_cast<size_t>(p.out_side) * 16;
    constexpr int KW = POOL == 31 ? 3 : 4;        // pooling window
    constexpr int LAG = POOL == 42 ? 2 : KW - 1;  // conv rows between an output row's first row and its last

    float bias_r[CG];
#pragma unroll
    for (int c = 0; c < CG; ++c) bias_r[c] = s_bias[grp * CG + c];
    // fused join: per-channel coefficients of this thread's channels and the residual tensor's strides
    constexpr bool kJoinPreload = JOIN && CG == 8 && POOL == 41 && !SPLIT;
    constexpr bool kJoinAFirst = true;  // A on the fp32 window sums (false: after the 16-bit rounding; same accuracy, more work)
    const uint32_t jplane_bytes = static_cast<uint32_t>(p.res_side) * 16;
    const uint32_t jrow_bytes = static_cast<uint32_t>(out_planes) * jplane_bytes;
    // every accumulator slot starts out holding the bias: the MMAs then always accumulate
    for (int s = 0; s < R; ++s) tc_st<CG>(t_base + s * COUT, bias_r);
    tc_wait_st();
    tc_fence_before();
    __syncwarp();
    if (lane == 0)
      for (int s = 0; s < RP; ++s) mbar_arrive(bar_acce0 + 8u * s);
    pdl_wait();  // before the first global store / residual gather

    uint32_t G = 0, iter = 0;
    for (int cur = row_lo; cur < row_hi;) {
      const Item it = decode_piece<POOL, SEG>(p, cur, row_hi);
      cur += it.npo;
      int n_img, col;
      bool col_ok;
      if (POOL == 0) {
        n_img = it.n0 + seg;
        col = it.x_in0 + xs;
        col_ok = col < p.conv_side;
      } else {
        // window `win` of this image starts win_step_out output columns after the previous one
        const int win = SEG == 1 ? quad : (quad & 1);
        n_img = it.n0 + (SEG == 1 ? 0 : (quad >> 1));
        const int lcol = POOL == 42 ? (lane >> 1) : lane;  // output column inside the window
        col = it.x_out0 + win * p.win_step_out + lcol;
        col_ok = lcol < p.win_step_out && !(POOL == 42 && (lane & 1)) && col < p.out_side;
      }
      col_ok = col_ok && n_img < p.N;
      // fused residual join: horizontal taps of this thread's output column (fixed for the item).  jsrc points at
      // the left tap of source row 0, jdx is the byte step to the right tap; jbot keeps the horizontally
      // interpolated lower source row of the previous output row, which is the upper one of the next output row
      // whenever the source rows advance by one (the 215 -> 205 join)
      const uint8_t* jsrc = nullptr;
      uint32_t jdx = 0, jtx2 = 0, jbot[JOIN ? NP : 1];
      int jy_prev = -1;
      uint4 jpl[2], jpr[2];  // kJoinPreload: left/right taps of the lower source row of the next two output rows
      float jtxf = 0.f, jbot_f[(JOIN && SPLIT) ? 8 : 1];  // split joins: fp32 horizontal weight, cached lower source row
      int jy_split = -2;
      if (JOIN) {
        const float fx = static_cast<float>(col) * p.res_scale;
        const int jx0 = static_cast<int>(fx);
        jdx = jx0 + 1 < p.res_side ? 16u : 0u;
        jtx2 = HH::splat(fx - static_cast<float>(jx0));
        jtxf = fx - static_cast<float>(jx0);
        jsrc = p.res_src + static_cast<size_t>(n_img) * p.res_side * out_planes * p.res_side * 16 +
               (static_cast<size_t>(part) * (CREAL / 8) + grp * (CG / 8)) * p.res_side * 16 + jx0 * 16;
      }
      // issue the gathers for output rows first_row, first_row + 1 (relative to the item); they are consumed one
      // loop iteration later, so their latency hides behind the accumulator wait and the pooling arithmetic
      auto join_preload = [&](int first_row) {
        if (col_ok) {
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            const int row = min(max(first_row + k, 0), it.npo - 1);
            const int y1 = min(static_cast<int>(static_cast<float>(it.po0 + row) * p.res_scale) + 1, p.res_side - 1);
            const uint8_t* a = jsrc + static_cast<uint32_t>(y1) * jrow_bytes;
            jpl[k] = *reinterpret_cast<const uint4*>(a);
            jpr[k] = *reinterpret_cast<const uint4*>(a + jdx);
          }
          // ... and pull the rows of the iteration after that towards L1/L2 (two source rows further down), so that
          // the register loads above rarely see a DRAM round trip
          const int row2 = min(first_row + 3, it.npo - 1);
          const int y2 = min(static_cast<int>(static_cast<float>(it.po0 + max(row2, 0)) * p.res_scale) + 1, p.res_side - 1);
          const uint8_t* a2 = jsrc + static_cast<uint32_t>(y2) * jrow_bytes;
          asm volatile("prefetch.global.L1 [%0];" ::"l"(a2));
          asm volatile("prefetch.global.L1 [%0];" ::"l"(a2 - jrow_bytes));
        }
      };
      if constexpr (kJoinPreload) join_preload(-LAG);
      // output row pointer of the row produced by output slot 0 of the current iteration (may start "before" po0)
      uint8_t* optr = p.out + n_img * out_img_bytes +
                      ((static_cast<size_t>(part) * (CREAL / 8) + grp * (CG / 8)) * p.out_side + col) * 16 +
                      static_cast<ptrdiff_t>(it.po0) * out_row_bytes + (CG == 4 ? grp * 8 : 0);
      if (POOL == 41 || POOL == 31) optr -= static_cast<ptrdiff_t>(LAG) * out_row_bytes;
      if (POOL == 42) optr -= out_row_bytes;

      // vertical pooling window: packed fp32 pairs (add.f32x2 - one issue slot per two channels)
      f32x2_t r1[(POOL == 41 || POOL == 31) ? NP : 1], q1[POOL != 0 ? NP : 1], q2[POOL == 41 ? NP : 1];
#pragma unroll
      for (int c = 0; c < (POOL != 0 ? NP : 1); ++c) q1[c] = 0ull;
#pragma unroll
      for (int c = 0; c < ((POOL == 41 || POOL == 31) ? NP : 1); ++c) r1[c] = 0ull;
#pragma unroll
      for (int c = 0; c < (POOL == 41 ? NP : 1); ++c) q2[c] = 0ull;

      // it.nconv is even: two conv rows per iteration (row slots gy, gy+1)
#pragma unroll 2
      for (int y = 0; y < it.nconv; y += 2, ++iter) {
        const uint32_t gy = G + y;
        const uint32_t slot0 = gy & (R - 1), slot1 = (gy + 1) & (R - 1);
        const uint32_t pair = (gy >> 1) & (RP - 1);
        mbar_wait(bar_accf0 + 8u * pair, (gy >> LOGR) & 1);
        tc_fence_after();
        if (RN_TC_DBG_BIT(p, 8)) {  // timing experiment: hand the slots straight back
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_acce0 + 8u * pair);
          optr += (POOL == 42 ? 1 : 2) * out_row_bytes;
          continue;
        }
        float a[CG], b[CG];
        tc_ld<CG>(t_base + slot0 * COUT, a);
        tc_ld<CG>(t_base + slot1 * COUT, b);
        tc_wait_ld();
        // hand the slots back, pre-loaded with the bias
        if (!RN_TC_DBG_BIT(p, 4)) {
          tc_st<CG>(t_base + slot0 * COUT, bias_r);
          tc_st<CG>(t_base + slot1 * COUT, bias_r);
          tc_wait_st();
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_acce0 + 8u * pair);

        // ---- saturate (= ReLU6/6) + vertical window -> packed 16-bit pairs vp[k][i] for output slot k
        // output slot 0 is pooled row (y - LAG) [41/31] or (y - 2)/2 [42] or conv row y [0]; slot 1 the next one
        uint32_t vp[2][NP];
        float vf[2][SPLIT ? CG : 1];  // split layers: the window sums stay fp32
#pragma unroll
        for (int i = 0; i < NP; ++i) {
          // weights and bias carry a factor 1/6: relu6(z)/6 == saturate(z/6), one FADD.SAT
          const f32x2_t x0 = f2_pack(__saturatef(a[2 * i]), __saturatef(a[2 * i + 1]));
          const f32x2_t x1 = f2_pack(__saturatef(b[2 * i]), __saturatef(b[2 * i + 1]));
          f32x2_t o0, o1 = 0ull;
          if (POOL == 0) {
            o0 = x0;
            o1 = x1;
          } else if (POOL == 42) {
            const f32x2_t q0 = f2_add(x0, x1);
            o0 = f2_add(q0, q1[i]);
            q1[i] = q0;
          } else if (POOL == 41) {  // sums of row pairs: q2 = p(y-3), q1 = p(y-2), r1 = x(y-1)
            const f32x2_t qa = f2_add(x0, r1[i]), qb = f2_add(x1, x0);
            o0 = f2_add(qa, q2[i]);
            o1 = f2_add(qb, q1[i]);
            q2[i] = qa;
            q1[i] = qb;
            r1[i] = x1;
          } else {  // 31: q1 = r(y-1) + r(y-2), r1 = r(y-1)
            o0 = f2_add(x0, q1[i]);
            o1 = f2_add(x1, f2_add(x0, r1[i]));
            q1[i] = f2_add(x1, x0);
            r1[i] = x1;
          }
          if constexpr (JOIN && kJoinAFirst && !SPLIT) {  // join coefficient A on the fp32 window sums (exact; the 16-bit sums then run on A*x)
            const f32x2_t a2 = *reinterpret_cast<const f32x2_t*>(s_abc + grp * CG + 2 * i);
            o0 = f2_mul(o0, a2);
            if (POOL != 42) o1 = f2_mul(o1, a2);
          }
          if constexpr (JOIN && SPLIT) {  // (kJoinAFirst applies to the 16-bit path only)
            const f32x2_t a2 = *reinterpret_cast<const f32x2_t*>(s_abc + grp * CG + 2 * i);
            o0 = f2_mul(o0, a2);
            if (POOL != 42) o1 = f2_mul(o1, a2);
          }
          float lo, hi;
          f2_unpack(o0, lo, hi);
          if constexpr (SPLIT) {
            vf[0][2 * i] = lo;
            vf[0][2 * i + 1] = hi;
          } else {
            vp[0][i] = HH::pack(lo, hi);
          }
          if (POOL != 42) {
            f2_unpack(o1, lo, hi);
            if constexpr (SPLIT) {
              vf[1][2 * i] = lo;
              vf[1][2 * i + 1] = hi;
            } else {
              vp[1][i] = HH::pack(lo, hi);
            }
          }
        }

        // ---- horizontal window with warp shuffles (every window owns its halo, see kWindows)
        uint32_t hp[2][NP];
        constexpr int NK = (POOL == 42) ? 1 : 2;
        if constexpr (SPLIT) {
#pragma unroll
          for (int k = 0; k < NK; ++k)
#pragma unroll
            for (int c = 0; c < CG; ++c) {
              const float v = vf[k][c];
              if (POOL == 42) {
                const float u = v + __shfl_xor_sync(0xffffffffu, v, 1);
                vf[k][c] = u + __shfl_down_sync(0xffffffffu, u, 2);
              } else if (POOL != 0) {
                const float t = v + __shfl_down_sync(0xffffffffu, v, 1);
                vf[k][c] = t + __shfl_down_sync(0xffffffffu, POOL == 41 ? t : v, 2);
              }
            }
        }
#pragma unroll
        for (int k = 0; k < (SPLIT ? 0 : NK); ++k)
#pragma unroll
          for (int i = 0; i < NP; ++i) {
            const uint32_t v = vp[k][i];
            if (POOL == 0 || RN_TC_DBG_BIT(p, 2)) {
              hp[k][i] = v;
            } else if (POOL == 42) {
              const uint32_t u = HH::add(v, __shfl_xor_sync(0xffffffffu, v, 1));
              hp[k][i] = HH::add(u, __shfl_down_sync(0xffffffffu, u, 2));
            } else {
              const uint32_t t = HH::add(v, __shfl_down_sync(0xffffffffu, v, 1));
              hp[k][i] = HH::add(t, __shfl_down_sync(0xffffffffu, POOL == 41 ? t : v, 2));
            }
          }
        // ---- stores: output slot k of this iteration is row (first + k) relative to the item
        {
          int first;  // index (relative to po0 / c0) of output slot 0
          if (POOL == 0) first = y;
          else if (POOL == 42) first = (y >> 1) - 1;
          else first = y - LAG;
          if (col_ok && !RN_TC_DBG_BIT(p, 1)) {
#pragma unroll
            for (int k = 0; k < NK; ++k) {
              const int row = first + k;
              if (row >= 0 && row < it.npo) {
                uint8_t* orow = optr + static_cast<size_t>(k) * out_row_bytes;
                if constexpr (SPLIT) {
                  // fp32-class path: residual join in fp32 from the hi + lo planes of the source tensor (reference
                  // network.py:199-203, TF-1.13 legacy bilinear: top = tl + (tr - tl) * tx, ...), then the hi / lo split
                  float ty = 0.f;
                  const uint8_t* r0 = nullptr;
                  const uint8_t* r1p = nullptr;
                  int jy0 = -1, jy1 = -1;
                  if constexpr (JOIN) {
                    const float fy = static_cast<float>(it.po0 + row) * p.res_scale;
                    jy0 = static_cast<int>(fy);
                    jy1 = min(jy0 + 1, p.res_side - 1);
                    ty = fy - static_cast<float>(jy0);
                    r0 = jsrc + static_cast<uint32_t>(jy0) * jrow_bytes;
                    r1p = jsrc + static_cast<uint32_t>(jy1) * jrow_bytes;
                  }
                  const size_t lo_off = static_cast<size_t>(p.cb_out_total) * out_plane_bytes;  // hi plane -> lo plane
                  const uint32_t jlo_off = static_cast<uint32_t>(p.cb_out_total) * jplane_bytes;
#pragma unroll
                  for (int cb = 0; cb < CG / 8; ++cb) {
                    float v[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) v[e] = vf[k][8 * cb + e];
                    if constexpr (JOIN) {
                      auto tap = [&](const uint8_t* rowp, uint32_t dxb, float* out8) {
                        const uint4 h = *reinterpret_cast<const uint4*>(rowp + cb * jplane_bytes + dxb);
                        const uint4 l = *reinterpret_cast<const uint4*>(rowp + cb * jplane_bytes + jlo_off + dxb);
                        const uint32_t hw[4] = {h.x, h.y, h.z, h.w}, lw[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                          const float2 a = HH::unpack(hw[q]), b = HH::unpack(lw[q]);
                          out8[2 * q] = a.x + b.x;
                          out8[2 * q + 1] = a.y + b.y;
                        }
                      };
                      // horizontal leg per source row; the lower row of this output row is the upper row of the next one
                      // whenever the source rows advance by one (conv2d_3's 215 -> 205 resize)
                      auto hrow = [&](const uint8_t* rowp, float* out8) {
                        float l8[8], r8[8];
                        tap(rowp, 0, l8);
                        tap(rowp, jdx, r8);
#pragma unroll
                        for (int e = 0; e < 8; ++e) out8[e] = l8[e] + (r8[e] - l8[e]) * jtxf;
                      };
                      float top[8], bot[8];
                      if (CG == 8 && jy_split == jy0) {
#pragma unroll
                        for (int e = 0; e < 8; ++e) top[e] = jbot_f[e];
                      } else {
                        hrow(r0, top);
                      }
                      hrow(r1p, bot);
                      if (CG == 8) {
#pragma unroll
                        for (int e = 0; e < 8; ++e) jbot_f[e] = bot[e];
                        jy_split = jy1;
                      }
                      {  // the next two output rows read the source rows below: start pulling them towards L1 now
                        const int step = POOL == 42 ? 2 : 1;
#pragma unroll
                        for (int a = 1; a <= 2; ++a) {
                          const uint8_t* pn = jsrc + static_cast<uint32_t>(min(jy1 + a * step, p.res_side - 1)) * jrow_bytes +
                                              cb * jplane_bytes;
                          asm volatile("prefetch.global.L1 [%0];" ::"l"(pn));
                          asm volatile("prefetch.global.L1 [%0];" ::"l"(pn + jlo_off));
                        }
                      }
#pragma unroll
                      for (int e = 0; e < 8; ++e) {
                        const float rs = top[e] + (bot[e] - top[e]) * ty;
                        const int ch = grp * CG + 8 * cb + e;
                        v[e] = v[e] + fmaf(s_abc[COUT + ch], rs, s_abc[2 * COUT + ch]);  // A was applied to the window sums
                      }
                    }
                    uint32_t hi4[4], lo4[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                      hi4[q] = HH::pack(v[2 * q], v[2 * q + 1]);
                      const float2 back = HH::unpack(hi4[q]);
                      lo4[q] = HH::pack(v[2 * q] - back.x, v[2 * q + 1] - back.y);
                    }
                    *reinterpret_cast<uint4*>(orow + cb * out_plane_bytes) = make_uint4(hi4[0], hi4[1], hi4[2], hi4[3]);
                    *reinterpret_cast<uint4*>(orow + cb * out_plane_bytes + lo_off) = make_uint4(lo4[0], lo4[1], lo4[2], lo4[3]);
                  }
                } else {
                if constexpr (JOIN) {
                  // reference network.py:199-203 in folded form; bilinear taps in
                  // packed 16-bit arithmetic (top = tl + (tr-tl)*tx, ...: the TF formula), the per-channel affine in
                  // fp32 (CPU emulation, DESIGN.md §2: the 16-bit lerp costs nothing measurable, a 16-bit affine
                  // would cost 4x the error budget)
                  const float fy = static_cast<float>(it.po0 + row) * p.res_scale;
                  const int y0 = static_cast<int>(fy);
                  const int y1 = min(y0 + 1, p.res_side - 1);
                  const uint32_t ty2 = HH::splat(fy - static_cast<float>(y0));
                  uint32_t top[NP];
                  if (y0 == jy_prev) {
#pragma unroll
                    for (int i = 0; i < NP; ++i) top[i] = jbot[i];
                  } else {
                    join_hlerp<HH, CG>(jsrc + static_cast<uint32_t>(y0) * jrow_bytes, jdx, jplane_bytes, jtx2, top);
                  }
                  if constexpr (kJoinPreload) {
                    jbot[0] = HH::fma(HH::sub(jpr[k].x, jpl[k].x), jtx2, jpl[k].x);
                    jbot[1] = HH::fma(HH::sub(jpr[k].y, jpl[k].y), jtx2, jpl[k].y);
                    jbot[2] = HH::fma(HH::sub(jpr[k].z, jpl[k].z), jtx2, jpl[k].z);
                    jbot[3] = HH::fma(HH::sub(jpr[k].w, jpl[k].w), jtx2, jpl[k].w);
                  } else {
                    join_hlerp<HH, CG>(jsrc + static_cast<uint32_t>(y1) * jrow_bytes, jdx, jplane_bytes, jtx2, jbot);
                    // the next output row reads source rows y1 + 1 (and y1 + 2): start pulling them into L1 now
                    const int yn = min(y1 + (POOL == 42 ? 2 : 1), p.res_side - 1);
                    const uint8_t* pn = jsrc + static_cast<uint32_t>(yn) * jrow_bytes;
#pragma unroll
                    for (int cb = 0; cb < CG / 8; ++cb)
                      asm volatile("prefetch.global.L1 [%0];" ::"l"(pn + cb * jplane_bytes));
                  }
                  jy_prev = y1;
#pragma unroll
                  for (int i = 0; i < NP; ++i) {
                    // out = A*pool (A already applied above) + (B * resized + C): B is a 16-bit value by construction
                    // (engine.cu stores the channel with a gain that makes it one), mixed-precision fma / add with
                    // fp32 accumulation straight from the packed 16-bit pairs
                    const uint32_t rs = HH::fma(HH::sub(jbot[i], top[i]), ty2, top[i]);
                    const uint32_t b2 = s_bh[(grp * CG) / 2 + i];
                    const float2 c2 = *reinterpret_cast<const float2*>(s_abc + 2 * COUT + grp * CG + 2 * i);
                    float lo = HH::fhfma_lo(rs, b2, c2.x), hi = HH::fhfma_hi(rs, b2, c2.y);
                    if constexpr (kJoinAFirst) {
                      lo = HH::fhadd_lo(hp[k][i], lo);
                      hi = HH::fhadd_hi(hp[k][i], hi);
                    } else {
                      const float2 hv = HH::unpack(hp[k][i]);
                      const float2 a2 = *reinterpret_cast<const float2*>(s_abc + grp * CG + 2 * i);
                      lo = fmaf(a2.x, hv.x, lo);
                      hi = fmaf(a2.y, hv.y, hi);
                    }
                    hp[k][i] = HH::pack(lo, hi);
                  }
                }
#pragma unroll
                for (int cb = 0; cb < CG / 8; ++cb)
                  *reinterpret_cast<uint4*>(orow + cb * out_plane_bytes) =
                      make_uint4(hp[k][4 * cb], hp[k][4 * cb + 1], hp[k][4 * cb + 2], hp[k][4 * cb + 3]);
                if (CG == 4) *reinterpret_cast<uint2*>(orow) = make_uint2(hp[k][0], hp[k][1]);
                }
              }
            }
          }
          optr += (POOL == 42 ? 1 : 2) * out_row_bytes;
          if constexpr (kJoinPreload) join_preload(first + NK);
        }
      }
      G += it.nconv;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(Cfg::kTmemCols) : "memory");
  }
}

// ---------------------------------------------------------------------------
// conv0: 3 -> 8 channels on CUDA cores in fp32 (K = 27 is too thin for the tensor
// pipe and its operands need more than 11 bits), + ReLU6 + 3x3/1
// average pool, output in chunked 16-bit layout (one chunk).
// Block = 32x16 pooled pixels; conv+ReLU6 staged in shared memory, then pooled.
// ---------------------------------------------------------------------------
constexpr int kC0W = 32, kC0H = 16;

template <typename TIn>
__global__ void __launch_bounds__(256) conv0_pool_kernel(const TIn* __restrict__ in, const float* __restrict__ w,
                                                         const float* __restrict__ bias, uint8_t* __restrict__ out, int S,
                                                         int bf16) {
  __shared__ float s_in[kC0H + 4][kC0W + 4][3];
  __shared__ float s_conv[kC0H + 2][kC0W + 2][8];
  __shared__ float s_w[27][8];
  __shared__ float s_b[8];
  const int PS = S - 4;  // pooled side
  const int n = blockIdx.z;
  const int y0 = blockIdx.y * kC0H, x0 = blockIdx.x * kC0W;
  const TIn* in_n = in + static_cast<size_t>(n) * S * S * 3;
  for (int i = threadIdx.x; i < 27 * 8; i += 256) s_w[i / 8][i % 8] = w[i];
  if (threadIdx.x < 8) s_b[threadIdx.x] = bias[threadIdx.x];
  for (int i = threadIdx.x; i < (kC0H + 4) * (kC0W + 4) * 3; i += 256) {
    int c = i % 3, px = (i / 3) % (kC0W + 4), py = i / (3 * (kC0W + 4));
    int iy = y0 + py, ix = x0 + px;
    s_in[py][px][c] = (iy < S && ix < S) ? static_cast<float>(in_n[(static_cast<size_t>(iy) * S + ix) * 3 + c]) : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < (kC0H + 2) * (kC0W + 2); i += 256) {
    int px = i % (kC0W + 2), py = i / (kC0W + 2);
    float acc[8];
#pragma unroll
    for (int o = 0; o < 8; ++o) acc[o] = 0.f;
#pragma unroll
    for (int tap = 0; tap < 9; ++tap)
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        float a = s_in[py + tap / 3][px + tap % 3][c];
#pragma unroll
        for (int o = 0; o < 8; ++o) acc[o] = fmaf(a, s_w[tap * 3 + c][o], acc[o]);
      }
#pragma unroll
    for (int o = 0; o < 8; ++o) s_conv[py][px][o] = relu6f(acc[o] + s_b[o]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kC0H * kC0W; i += 256) {
    int px = i % kC0W, py = i / kC0W;
    int oy = y0 + py, ox = x0 + px;
    if (oy >= PS || ox >= PS) continue;
    float h[8];
#pragma unroll
    for (int o = 0; o < 8; ++o) {
      float s = 0.f;
#pragma unroll
      for (int dy = 0; dy < 3; ++dy)
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) s += s_conv[py + dy][px + dx][o];
      h[o] = s * (1.f / 6.f);  // stored scale 9/6 of the pooled mean, as the tensor-core layers (engine.cu)
    }
    uint4 o4;
    o4.x = pack2(h[0], h[1], bf16);
    o4.y = pack2(h[2], h[3], bf16);
    o4.z = pack2(h[4], h[5], bf16);
    o4.w = pack2(h[6], h[7], bf16);
    *reinterpret_cast<uint4*>(out + ((static_cast<size_t>(n) * PS + oy) * PS + ox) * 16) = o4;
  }
}


// ---------------------------------------------------------------------------
// Fused tail for small maps (im_side 224: 21x21x16 in): conv 16->16 + ReLU6 + pool4/2, conv 16->16 + ReLU6 +
// pool4/2, residual join with the bilinearly resized block input, 4 dense layers, softmax, argmax
// (reference network.py:230-237, :44-45 in folded form).  One CTA per image, everything in shared memory, fp32.
// ---------------------------------------------------------------------------
struct TailParams {
  const float* w8;  // [9][16][16] HWIO folded
  const float* b8;
  const float* w9;
  const float* b9;
  const float* ja;  // join coefficients (per channel) for TRUE-scale pooled tensors
  const float* jb;
  const float* jc;
  DenseParams dense;
  int S7;           // side of the input map
  float in_scale;   // stored -> true
  int flat_len;
  int split;        // the input tensor carries hi + lo planes (RN_PREC_FP32_TC)
};

constexpr int kTailC = 16;
constexpr int kTailMaxSide = 24;

__device__ __forceinline__ void tail_conv_pool(const float* __restrict__ in, int S, const float* __restrict__ w,
                                               const float* __restrict__ bias, float* __restrict__ conv,
                                               float* __restrict__ pooled, const float* __restrict__ s_w) {
  // in: [16][S][S] channel-major; conv: [16][S-2][S-2]; pooled: [16][P][P], P = (S-2-4)/2+1; s_w: the layer's
  // [9][16][16] weights, already in shared memory (`w` unused: kept for the signature of the callers' tables)
  (void)w;
  const int CS = S - 2, P = (CS - 4) / 2 + 1;
  // One item = two horizontally adjacent pixels x eight output channels: the four input values of a kernel row serve
  // both pixels, the weights of a tap come in as two 128-bit loads, and the 16 accumulators advance as packed fp32 pairs
  // (fma.f32x2).  Per accumulator the products are added in the same order as before (channel, then dy, dx).
  const int XP = (CS + 1) / 2;
  for (int item = threadIdx.x; item < CS * XP * 2; item += blockDim.x) {
    const int g = item / (CS * XP), r = item % (CS * XP);
    const int y = r / XP, x = (r % XP) * 2;
    const bool two = x + 1 < CS;
    f32x2_t acc[2][4];
#pragma unroll
    for (int k = 0; k < 4; ++k) acc[0][k] = acc[1][k] = 0ull;
    const int x1 = min(x + 1, S - 1), x2 = min(x + 2, S - 1), x3 = min(x + 3, S - 1);
    for (int c = 0; c < 16; ++c) {
#pragma unroll
      for (int dy = 0; dy < 3; ++dy) {
        const float* ip = in + (c * S + y + dy) * S;
        const float a0 = ip[x], a1 = ip[x1], a2 = ip[x2], a3 = ip[x3];
        const f32x2_t av[4] = {f2_pack(a0, a0), f2_pack(a1, a1), f2_pack(a2, a2), f2_pack(a3, a3)};
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
          const float* wp = &s_w[((dy * 3 + dx) * 16 + c) * 16 + g * 8];
          const ulonglong2 w01 = *reinterpret_cast<const ulonglong2*>(wp);
          const ulonglong2 w23 = *reinterpret_cast<const ulonglong2*>(wp + 4);
          acc[0][0] = f2_fma(av[dx], w01.x, acc[0][0]);
          acc[0][1] = f2_fma(av[dx], w01.y, acc[0][1]);
          acc[0][2] = f2_fma(av[dx], w23.x, acc[0][2]);
          acc[0][3] = f2_fma(av[dx], w23.y, acc[0][3]);
          acc[1][0] = f2_fma(av[dx + 1], w01.x, acc[1][0]);
          acc[1][1] = f2_fma(av[dx + 1], w01.y, acc[1][1]);
          acc[1][2] = f2_fma(av[dx + 1], w23.x, acc[1][2]);
          acc[1][3] = f2_fma(av[dx + 1], w23.y, acc[1][3]);
        }
      }
    }
#pragma unroll
    for (int px = 0; px < 2; ++px) {
      if (px == 1 && !two) break;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float lo, hi;
        f2_unpack(acc[px][k], lo, hi);
        conv[((g * 8 + 2 * k) * CS + y) * CS + x + px] = relu6f(lo + bias[g * 8 + 2 * k]);
        conv[((g * 8 + 2 * k + 1) * CS + y) * CS + x + px] = relu6f(hi + bias[g * 8 + 2 * k + 1]);
      }
    }
  }
  __syncthreads();
  for (int item = threadIdx.x; item < 16 * P * P; item += blockDim.x) {
    const int c = item / (P * P), py = (item / P) % P, qx = item % P;
    float sum = 0.f;
    for (int dy = 0; dy < 4; ++dy)
      for (int dx = 0; dx < 4; ++dx) sum += conv[(c * CS + 2 * py + dy) * CS + 2 * qx + dx];
    pooled[item] = sum * 0.0625f;
  }
  __syncthreads();
}

__global__ void __launch_bounds__(1024) tail_fused_kernel(const uint16_t* __restrict__ p7, TailParams tp, int bf16,
                                                         long long* __restrict__ top1, float* __restrict__ probs,
                                                         float* __restrict__ logits, float* __restrict__ dbg8,
                                                         float* __restrict__ dbg9) {
  extern __shared__ __align__(16) float sm[];
  const int S = tp.S7, S8 = (S - 2 - 4) / 2 + 1, S9 = (S8 - 2 - 4) / 2 + 1;
  float* s_in = sm;                                  // [16][S][S]
  float* s_conv = s_in + 16 * S * S;                 // [16][S-2][S-2]
  float* s_p8 = s_conv + 16 * (S - 2) * (S - 2);     // [16][S8][S8]
  float* s_p9 = s_p8 + 16 * S8 * S8;                 // [16][S9][S9]
  float* s_flat = s_p9 + 16 * S9 * S9;               // [S9*S9*16] NHWC order
  float* s_w8 = s_flat + 16 * S9 * S9;               // [9*16*16]
  float* s_w9 = s_w8 + 9 * 16 * 16;                  // [9*16*16]
  float* s_dw = s_w9 + 9 * 16 * 16;                  // dense kernels [in][out], layer after layer, then the biases
  const DenseParams& dp = tp.dense;
  const int n = blockIdx.x;
  // Everything the image needs comes in with ONE round of global loads (the input map, both conv kernels and the
  // dense head): at batch 1 this kernel is pure latency, and every later phase then runs out of shared memory.
  for (int i = threadIdx.x; i < 9 * 16 * 16; i += blockDim.x) {
    s_w8[i] = tp.w8[i];
    s_w9[i] = tp.w9[i];
  }
  int dw_off[5];
  dw_off[0] = 0;
#pragma unroll
  for (int l = 0; l < 4; ++l) {
    const int cnt = (l == 0 ? tp.flat_len : dp.out[l - 1]) * dp.out[l];
    for (int i = threadIdx.x; i < cnt; i += blockDim.x) s_dw[dw_off[l] + i] = dp.w[l][i];
    dw_off[l + 1] = dw_off[l] + cnt;
  }
  float* s_db = s_dw + dw_off[4];
  int db_off[4];
  {
    int o = 0;
#pragma unroll
    for (int l = 0; l < 4; ++l) {
      db_off[l] = o;
      for (int i = threadIdx.x; i < dp.out[l]; i += blockDim.x) s_db[o + i] = dp.b[l][i];
      o += dp.out[l];
    }
  }
  pdl_wait();  // everything above is constant data; p7 is the previous kernel's output
  // chunked 16-bit [y][cb(2)][x][8] -> channel-major fp32, true scale
  // (fp32-class path: rows of [hi0 | hi1 | lo0 | lo1] planes, the value is hi + lo)
  const int planes = tp.split ? 4 : 2;
  const uint16_t* src = p7 + static_cast<size_t>(n) * S * S * 8 * planes;
  for (int i = threadIdx.x; i < S * S * kTailC; i += blockDim.x) {
    const int e = i % 8, x = (i / 8) % S, cb = (i / (8 * S)) % 2, y = i / (16 * S);
    const int at = ((y * planes + cb) * S + x) * 8 + e;
    const uint16_t raw = src[at];
    float f;
    const uint16_t raw_lo = tp.split ? src[at + 2 * S * 8] : static_cast<uint16_t>(0);
    if (bf16) {
      f = __uint_as_float(static_cast<uint32_t>(raw) << 16) + __uint_as_float(static_cast<uint32_t>(raw_lo) << 16);
    } else {
      f = __half2float(*reinterpret_cast<const __half*>(&raw)) + __half2float(*reinterpret_cast<const __half*>(&raw_lo));
    }
    s_in[((cb * 8 + e) * S + y) * S + x] = f * tp.in_scale;
  }
  __syncthreads();
  tail_conv_pool(s_in, S, tp.w8, tp.b8, s_conv, s_p8, s_w8);
  tail_conv_pool(s_p8, S8, tp.w9, tp.b9, s_conv, s_p9, s_w9);
  if (dbg8)
    for (int i = threadIdx.x; i < 16 * S8 * S8; i += blockDim.x) {  // NHWC for the parity tests
      const int c = i % 16, px = i / 16;
      dbg8[static_cast<size_t>(n) * 16 * S8 * S8 + i] = s_p8[c * S8 * S8 + px];
    }
  // residual join: A*p9 + B*resize(p7 -> S9 x S9) + C, written in NHWC flatten order (h, w, c)
  const float scale = static_cast<float>(S) / static_cast<float>(S9);
  for (int i = threadIdx.x; i < 16 * S9 * S9; i += blockDim.x) {
    const int c = i % 16, x = (i / 16) % S9, y = i / (16 * S9);
    const float fy = static_cast<float>(y) * scale, fx = static_cast<float>(x) * scale;
    const int y0 = static_cast<int>(floorf(fy)), x0 = static_cast<int>(floorf(fx));
    const int y1 = min(y0 + 1, S - 1), x1 = min(x0 + 1, S - 1);
    const float ty = fy - static_cast<float>(y0), tx = fx - static_cast<float>(x0);
    const float* ch = s_in + c * S * S;
    const float tl = ch[y0 * S + x0], tr = ch[y0 * S + x1], bl = ch[y1 * S + x0], br = ch[y1 * S + x1];
    const float top = tl + (tr - tl) * tx, bot = bl + (br - bl) * tx;
    const float rs = top + (bot - top) * ty;
    const float v = fmaf(tp.ja[c], s_p9[(c * S9 + y) * S9 + x], fmaf(tp.jb[c], rs, tp.jc[c]));
    s_flat[i] = v;
    if (dbg9) dbg9[static_cast<size_t>(n) * 16 * S9 * S9 + i] = v;
  }
  __syncthreads();
  // dense head on warp 0 (same arithmetic as dense_tail_kernel in kernels_f32.cu)
  if (threadIdx.x >= 32) return;
  const int lane = threadIdx.x;
  float v = 0.f;
  {
    const int on = dp.out[0];
    if (lane < on) {
      float acc = 0.f;
      for (int r = 0; r < tp.flat_len; ++r) acc = fmaf(s_flat[r], s_dw[r * on + lane], acc);
      v = relu6f(acc + s_db[db_off[0] + lane]);
    }
  }
#pragma unroll
  for (int l = 1; l < 4; ++l) {
    const int in = dp.out[l - 1], on = dp.out[l];
    const float* wl = s_dw + dw_off[l];
    float acc = 0.f;
    for (int r = 0; r < in; ++r) {
      float xr = __shfl_sync(0xffffffffu, v, r);
      if (lane < on) acc = fmaf(xr, wl[r * on + lane], acc);
    }
    if (lane < on) acc += s_db[db_off[l] + lane];
    v = relu6f(acc);
  }
  const int C = dp.out[3];
  float m = lane < C ? v : -INFINITY;
  for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  float e = lane < C ? expf(v - m) : 0.f;
  float ssum = e;
  for (int o = 16; o; o >>= 1) ssum += __shfl_xor_sync(0xffffffffu, ssum, o);
  const float prob = e / ssum;
  float best = lane < C ? prob : -1.f;
  int bi = lane;
  for (int o = 16; o; o >>= 1) {
    float ob = __shfl_xor_sync(0xffffffffu, best, o);
    int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ob > best || (ob == best && oi < bi)) {
      best = ob;
      bi = oi;
    }
  }
  if (lane < C) {
    if (probs) probs[static_cast<size_t>(n) * C + lane] = prob;
    if (logits) logits[static_cast<size_t>(n) * C + lane] = v;
  }
  if (lane == 0 && top1) top1[n] = bi;
}

// `split`: the tensor carries hi planes followed by lo planes (fp32-class path); the value is hi + lo
__global__ void chunked_to_f32_kernel(const uint16_t* __restrict__ in, float* __restrict__ out, int N, int S, int C,
                                      int bf16, float scale, int split) {
  const size_t total = static_cast<size_t>(N) * S * S * C;
  const int CBl = C / 8, CBn = split ? 2 * CBl : CBl;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    int c = static_cast<int>(i % C);
    size_t r = i / C;
    int x = static_cast<int>(r % S);
    r /= S;
    int y = static_cast<int>(r % S);
    int n = static_cast<int>(r / S);
    const size_t at = ((((static_cast<size_t>(n) * S + y) * CBn + c / 8) * S) + x) * 8 + (c % 8);
    uint16_t raw = in[at];
    float f;
    if (bf16) {
      f = __uint_as_float(static_cast<uint32_t>(raw) << 16);
    } else {
      __half h = *reinterpret_cast<__half*>(&raw);
      f = __half2float(h);
    }
    if (split) {
      uint16_t raw_lo = in[at + static_cast<size_t>(CBl) * S * 8];
      f += bf16 ? __uint_as_float(static_cast<uint32_t>(raw_lo) << 16) : __half2float(*reinterpret_cast<__half*>(&raw_lo));
    }
    out[i] = f * scale;
  }
}

// uint8 NHWC (3 channels) -> conv0 tensor-core input: out[n][y][x] = 16 bytes {p(x).c0,c1,c2,0, p(x+1).c0,c1,c2,0}
// with p = pixel/256 (exact in fp16 and bf16).  One thread per output chunk.
__global__ void prep_u8_kernel(const uint8_t* __restrict__ in, uint4* __restrict__ out, int N, int S, int bf16,
                               int px_bytes) {
  const size_t total = static_cast<size_t>(N) * S * S;
  pdl_trigger();
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int x = static_cast<int>(i % S);
    const uint8_t* px = in + i * px_bytes;  // 3 = packed BGR/RGB, 4 = BGRA (Android ARGB_8888 ints)
    float a0 = px[0], a1 = px[1], a2 = px[2];
    float b0 = 0.f, b1 = 0.f, b2 = 0.f;
    if (x + 1 < S) {
      b0 = px[px_bytes];
      b1 = px[px_bytes + 1];
      b2 = px[px_bytes + 2];
    }
    const float sc = 1.f / 256.f;
    uint4 o;
    o.x = pack2(a0 * sc, a1 * sc, bf16);
    o.y = pack2(a2 * sc, 0.f, bf16);
    o.z = pack2(b0 * sc, b1 * sc, bf16);
    o.w = pack2(b2 * sc, 0.f, bf16);
    out[i] = o;
  }
}

// Fast form for packed 3-byte pixels, fp16 and a 4-byte aligned image pointer with S * 3 % 4 == 0: one thread per output
// chunk as above (coalesced 16-byte stores), but the six bytes come from three aligned 32-bit loads + funnel shifts and
// are converted with byte permutes and packed fp16 math instead of six byte loads and float conversions.
// 0x6400 | b is the fp16 number 1024 + b, so (0x6400 | b) * 2^-8 - 4 = b / 256 exactly.
__global__ void prep_u8w_kernel(const uint32_t* __restrict__ in, uint4* __restrict__ out, size_t total_px, int S) {
  pdl_trigger();
  const uint32_t kBias = 0x64006400u;  // bytes {0x00, 0x64, 0x00, 0x64}: selector 4 = 0x00, 5 = 0x64
  const __half2 scale = __floats2half2_rn(1.f / 256.f, 1.f / 256.f), minus4 = __floats2half2_rn(-4.f, -4.f);
  auto cvt = [&](uint32_t q, uint32_t sel) {
    const uint32_t v = __byte_perm(q, kBias, sel);
    const __half2 r = __hfma2(*reinterpret_cast<const __half2*>(&v), scale, minus4);
    return *reinterpret_cast<const uint32_t*>(&r);
  };
  const size_t last_word = (total_px * 3 + 3) / 4 - 1;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total_px;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t o = 3 * i, k = o >> 2;
    const uint32_t sh = static_cast<uint32_t>(o & 3) * 8;
    const uint32_t w0 = in[k], w1 = in[min(k + 1, last_word)], w2 = in[min(k + 2, last_word)];
    const uint32_t a = __funnelshift_r(w0, w1, sh);  // bytes o .. o+3
    const uint32_t b = __funnelshift_r(w1, w2, sh);  // bytes o+4 .. o+7
    uint32_t q1 = __byte_perm(a, b, 0x0543);         // pixel x+1 = bytes o+3 .. o+5
    if (static_cast<int>(i % S) + 1 >= S) q1 = 0u;   // the row ends here: no right neighbour
    out[i] = make_uint4(cvt(a, 0x5150), cvt(a, 0x5452), cvt(q1, 0x5150), cvt(q1, 0x5452));
  }
}

template <int CB, int COUT, int POOL, int SEG, int AMODE, bool BF16, int CREAL = COUT, bool JOIN = false, bool SPLIT = false>
cudaError_t launch_tc_impl(const TcConvLayer& L, const void* in, void* out, int N, cudaStream_t st) {
  using Cfg = TcCfg<CB, COUT, AMODE, POOL != 0, SPLIT && AMODE == 0>;
  TcParams p{};
  p.in = static_cast<const uint8_t*>(in);
  p.out = static_cast<uint8_t*>(out);
  p.w = static_cast<const uint8_t*>(L.w_packed);
  p.bias = L.bias;
  p.N = N;
  p.in_side = L.in_side;
  p.conv_side = L.in_side - 2;
  p.out_side = L.out_side;
  p.cb_out_total = (CREAL * L.cout_parts) / 8;
  p.w_bytes = static_cast<int>(L.w_bytes);
  constexpr int SEGW = kTileM / SEG;
  if (SEG == 2 && L.in_side > SEGW) return cudaErrorInvalidValue;
  if (POOL == 0) {
    p.strip_step_in = p.strip_step_out = kTileM;
    p.n_strips = SEG == 2 ? 1 : (p.conv_side + kTileM - 1) / kTileM;
  } else {
    // a 32-lane window yields 30 valid conv columns (taps of lanes 30/31 cross into the next window)
    p.win_step_out = POOL == 41 ? 27 : (POOL == 31 ? 28 : 14);
    p.win_step_in = POOL == 42 ? 2 * p.win_step_out : p.win_step_out;
    const int wins = SEG == 2 ? 2 : 4;  // SEG == 2: the tile is 2 windows x 2 images
    p.strip_step_out = wins * p.win_step_out;
    p.strip_step_in = wins * p.win_step_in;
    if (SEG == 2 && p.out_side > p.strip_step_out) return cudaErrorInvalidValue;
    p.n_strips = SEG == 2 ? 1 : (p.out_side + p.strip_step_out - 1) / p.strip_step_out;
  }
  const int groups = (N + SEG - 1) / SEG;
  // every gang of n_strips CTAs gets the same share of the (image group x output row) line (tc_common.cuh)
  p.total_rows = groups * p.out_side;
  int gangs = 1;
  rn_plan_rows(p.total_rows, p.out_side, p.n_strips, std::max(p.n_strips, SmCount() / L.cout_parts), &gangs, &p.min_piece);
  const int gx = gangs * p.n_strips;
  auto kern = conv_tc_kernel<CB, COUT, POOL, SEG, AMODE, BF16, CREAL, JOIN, SPLIT>;
  if (JOIN) {
    if (!L.join_src || !L.join_abc) return cudaErrorInvalidValue;
    p.res_src = static_cast<const uint8_t*>(L.join_src);
    p.res_side = L.join_src_side;
    p.res_scale = static_cast<float>(L.join_src_side) / static_cast<float>(L.out_side);
    p.join_abc = L.join_abc;
  }
  // per device (replicas of several GPUs share the process), and cheap enough to repeat
  cudaError_t ea = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
  if (ea != cudaSuccess) return ea;
  dim3 grid(gx, L.cout_parts);
  CUtensorMap tmap;
  std::memset(&tmap, 0, sizeof(tmap));
  if (POOL != 0) {
    // dims (fastest first): 256 elements = one 32-pixel window | window index (stride = win_step pixels, overlapping)
    //                       | channel-chunk plane | image row.  Box = all four windows x all planes of one row.
    const PFN_encodeTiled encode = GetEncodeTiled();
    if (!encode) return cudaErrorNotSupported;
    if (SEG == 2) {
      // {window elements | window | image | plane | row of the image}; box = 2 windows x 2 images x all planes x 2 rows,
      // which lands in shared memory as [row][plane][image][window]: lane quadrant = 2 * image + window
      const cuuint64_t row_bytes = static_cast<cuuint64_t>(CB) * p.in_side * 16;
      const cuuint64_t gdim5[5] = {256, 2, static_cast<cuuint64_t>(N), static_cast<cuuint64_t>(CB),
                                   static_cast<cuuint64_t>(p.in_side)};
      const cuuint64_t gstr5[4] = {static_cast<cuuint64_t>(p.win_step_in) * 16, row_bytes * p.in_side,
                                   static_cast<cuuint64_t>(p.in_side) * 16, row_bytes};
      const cuuint32_t box5[5] = {256, 2, 2, static_cast<cuuint32_t>(CB), 2};
      const cuuint32_t estr5[5] = {1, 1, 1, 1, 1};
      CUresult cr5 = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT16, 5, const_cast<uint8_t*>(p.in), gdim5, gstr5, box5, estr5,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (cr5 != CUDA_SUCCESS) return cudaErrorInvalidValue;
    } else {
    const cuuint64_t gdim[4] = {256, static_cast<cuuint64_t>(4 * p.n_strips), static_cast<cuuint64_t>(CB),
                                static_cast<cuuint64_t>(N) * p.in_side};
    const cuuint64_t gstr[3] = {static_cast<cuuint64_t>(p.win_step_in) * 16, static_cast<cuuint64_t>(p.in_side) * 16,
                                static_cast<cuuint64_t>(CB) * p.in_side * 16};
    const cuuint32_t box[4] = {256, 4, static_cast<cuuint32_t>(CB), 2};  // a pair of input rows per TMA
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult cr = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT16, 4, const_cast<uint8_t*>(p.in), gdim, gstr, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) return cudaErrorInvalidValue;
    }
  }
#ifdef RN_TC_TIMING_EXPERIMENTS  // never defined for the shipped library (see tools/experiments/README.md)
  if (const char* d = std::getenv("RN_TC_DBG")) p.dbg = std::atoi(d);
#endif
  cudaError_t el = LaunchPdl(kern, grid, dim3(tc_threads(CREAL)), Cfg::kSmemBytes, st, N, p, tmap);
  if (el != cudaSuccess) return el;
  return cudaGetLastError();
}

template <int CB, int COUT, int POOL, int SEG, int AMODE, int CREAL = COUT, bool JOIN = false>
cudaError_t launch_tc(const TcConvLayer& L, const void* in, void* out, int N, HalfKind kind, cudaStream_t st) {
  return kind == HalfKind::kBF16 ? launch_tc_impl<CB, COUT, POOL, SEG, AMODE, true, CREAL, JOIN>(L, in, out, N, st)
                                 : launch_tc_impl<CB, COUT, POOL, SEG, AMODE, false, CREAL, JOIN>(L, in, out, N, st);
}
// split layers: fp16 halves (RN_PREC_FP32_TC) or bf16 halves (RN_PREC_BF16X3)
template <int CB, int COUT, int POOL, int SEG, int AMODE, int CREAL = COUT, bool JOIN = false>
cudaError_t launch_tc_split(const TcConvLayer& L, const void* in, void* out, int N, HalfKind kind, cudaStream_t st) {
  return kind == HalfKind::kBF16 ? launch_tc_impl<CB, COUT, POOL, SEG, AMODE, true, CREAL, JOIN, true>(L, in, out, N, st)
                                 : launch_tc_impl<CB, COUT, POOL, SEG, AMODE, false, CREAL, JOIN, true>(L, in, out, N, st);
}

uint16_t to_half_bits(double v, HalfKind kind) {
  float f = static_cast<float>(v);
  if (kind == HalfKind::kBF16) {
    __nv_bfloat16 h = __float2bfloat16_rn(f);
    uint16_t u;
    std::memcpy(&u, &h, 2);
    return u;
  }
  __half h = __float2half_rn(f);
  uint16_t u;
  std::memcpy(&u, &h, 2);
  return u;
}

}  // namespace

size_t ChunkedBytes(int n, int side, int channels) {
  return static_cast<size_t>(n) * side * side * channels * 2 + kSlackBytes;
}

double RoundToHalfKind(double v, HalfKind kind) {
  const uint16_t bits = to_half_bits(v, kind);
  if (kind == HalfKind::kBF16) {
    const uint32_t u = static_cast<uint32_t>(bits) << 16;
    float f;
    std::memcpy(&f, &u, 4);
    return static_cast<double>(f);
  }
  __half_raw hr;
  hr.x = bits;
  return static_cast<double>(__half2float(__half(hr)));
}

size_t PackTcWeights(const double* w, int cin, int cout, int cout_parts, HalfKind kind, double scale, void* out_host,
                     const double* in_scale) {
  const int cb = cin / 8;
  const bool paired = cb == 1;
  const int planes = paired ? 4 : 3 * cb;
  const int cp = cout / cout_parts;
  const size_t part_bytes = static_cast<size_t>(planes) * 3 * cp * 16;
  if (!out_host) return part_bytes;
  uint16_t* o = static_cast<uint16_t*>(out_host);
  std::memset(o, 0, part_bytes * cout_parts);
  for (int part = 0; part < cout_parts; ++part)
    for (int pl = 0; pl < planes; ++pl) {
      const int dx = paired ? pl : pl / cb;
      const int c0 = paired ? 0 : (pl % cb) * 8;
      if (dx > 2) continue;  // zero plane that completes the K=16 pair of a Cin=8 layer
      for (int j = 0; j < 3; ++j) {
        const int dy = 2 - j;
        for (int oc = 0; oc < cp; ++oc)
          for (int e = 0; e < 8; ++e) {
            double v = scale * w[(((dy * 3 + dx) * cin) + c0 + e) * cout + part * cp + oc];
            if (in_scale) v *= in_scale[c0 + e];
            o[part * (part_bytes / 2) + ((static_cast<size_t>(pl) * 3 * cp + j * cp + oc) * 8) + e] =
                to_half_bits(v, kind);
          }
      }
    }
  return part_bytes;
}

size_t PackTcWeightsSplit(const double* w, int cin, int cout, int cout_parts, HalfKind kind, double scale, void* out_host) {
  // per part: [Wh image | Wl image], each laid out as PackTcWeights does; wh = round16(scale * w), wl = round16(rest)
  const size_t pb = PackTcWeights(nullptr, cin, cout, cout_parts, kind, 1.0, nullptr);
  if (!out_host) return 2 * pb;
  const size_t n = static_cast<size_t>(9) * cin * cout;
  std::vector<double> wh(n), wl(n);
  for (size_t i = 0; i < n; ++i) {
    const double v = scale * w[i];
    wh[i] = RoundToHalfKind(v, kind);
    wl[i] = RoundToHalfKind(v - wh[i], kind);
  }
  std::vector<uint8_t> hi(pb * cout_parts), lo(pb * cout_parts);
  PackTcWeights(wh.data(), cin, cout, cout_parts, kind, 1.0, hi.data());
  PackTcWeights(wl.data(), cin, cout, cout_parts, kind, 1.0, lo.data());
  uint8_t* o = static_cast<uint8_t*>(out_host);
  for (int part = 0; part < cout_parts; ++part) {
    std::memcpy(o + 2 * pb * part, hi.data() + pb * part, pb);
    std::memcpy(o + 2 * pb * part + pb, lo.data() + pb * part, pb);
  }
  return 2 * pb;
}

size_t PackTcConv0Weights(const double* w, HalfKind kind, double scale, void* out_host) {
  // planes: 0/1 = hi halves for pixel pairs (x,x+1)/(x+2,x+3), 2/3 = lo halves; rows [dy=2|dy=1|dy=0] x 16 (8 real)
  constexpr int kCout = 16, kReal = 8;
  const size_t bytes = static_cast<size_t>(4) * 3 * kCout * 16;
  if (!out_host) return bytes;
  uint16_t* o = static_cast<uint16_t*>(out_host);
  std::memset(o, 0, bytes);
  auto to_f = [&](uint16_t bits) {
    if (kind == HalfKind::kBF16) {
      uint32_t u = static_cast<uint32_t>(bits) << 16;
      float f;
      std::memcpy(&f, &u, 4);
      return static_cast<double>(f);
    }
    return static_cast<double>(__half2float(__ushort_as_half(bits)));
  };
  for (int half = 0; half < 2; ++half)
    for (int j = 0; j < 3; ++j)
      for (int oc = 0; oc < kReal; ++oc)
        for (int e = 0; e < 8; ++e) {
          const int ch = e % 4;
          if (ch == 3) continue;
          for (int hi_lo = 0; hi_lo < 2; ++hi_lo) {
            const int dx = half * 2 + e / 4;
            if (dx > 2) continue;
            const int dy = 2 - j;
            const double v = scale * w[((dy * 3 + dx) * 3 + ch) * kReal + oc];
            const uint16_t hi = to_half_bits(v, kind);
            const uint16_t bits = hi_lo == 0 ? hi : to_half_bits(v - to_f(hi), kind);
            const int pl = hi_lo * 2 + half;
            o[(static_cast<size_t>(pl) * 3 * kCout + j * kCout + oc) * 8 + e] = bits;
          }
        }
  return bytes;
}

cudaError_t PrepU8(const uint8_t* in, void* out, int N, int S, HalfKind kind, cudaStream_t st, int px_bytes) {
  size_t total = static_cast<size_t>(N) * S * S;
  if (px_bytes == 3 && kind == HalfKind::kF16 && (total * 3) % 4 == 0 && reinterpret_cast<uintptr_t>(in) % 4 == 0) {
    int blocksw = static_cast<int>(std::min<size_t>((total + 255) / 256, static_cast<size_t>(SmCount()) * 16));
    prep_u8w_kernel<<<blocksw, 256, 0, st>>>(reinterpret_cast<const uint32_t*>(in), static_cast<uint4*>(out), total, S);
    return cudaGetLastError();
  }
  int blocks = static_cast<int>(std::min<size_t>((total + 255) / 256, static_cast<size_t>(SmCount()) * 16));
  prep_u8_kernel<<<blocks, 256, 0, st>>>(in, static_cast<uint4*>(out), N, S, kind == HalfKind::kBF16, px_bytes);
  return cudaGetLastError();
}

cudaError_t ConvTc(const TcConvLayer& L, const void* in, void* out, int N, HalfKind kind, cudaStream_t st) {
  const int cb = L.cin / 8, cp = L.cout / L.cout_parts;
  // two images per tile: 2 x 64 lanes (no pooling) or 2 x 2 windows of 14 pooled columns (4x4/2 pooling)
  const bool seg2 = L.pool_k ? (L.out_side <= 28 && L.in_side <= 64) : L.in_side <= 64;
  const int pool = L.pool_k * 10 + L.pool_s;
  if (L.split) {  // fp32-class path: L.cin counts the physical (hi + lo) channels
    if (L.amode == 2) return launch_tc_split<1, 16, 31, 1, 2, 16>(L, in, out, N, kind, st);
    if (cb == 4 && cp == 32 && pool == 41 && !L.join_src) return launch_tc_split<4, 32, 41, 1, 0>(L, in, out, N, kind, st);
    if (cb == 8 && cp == 32 && pool == 41)
      return L.join_src ? launch_tc_split<8, 32, 41, 1, 0, 32, true>(L, in, out, N, kind, st)
                        : launch_tc_split<8, 32, 41, 1, 0>(L, in, out, N, kind, st);
    if (cb == 8 && cp == 64 && pool == 42 && !L.join_src) return launch_tc_split<8, 64, 42, 1, 0>(L, in, out, N, kind, st);
    if (cb == 16 && cp == 32 && pool == 42 && L.join_src) return launch_tc_split<16, 32, 42, 1, 0, 32, true>(L, in, out, N, kind, st);
    if (cb == 16 && cp == 32 && pool == 0 && !L.join_src)
      return seg2 ? launch_tc_split<16, 32, 0, 2, 0>(L, in, out, N, kind, st) : launch_tc_split<16, 32, 0, 1, 0>(L, in, out, N, kind, st);
    if (cb == 32 && cp == 16 && pool == 42 && !L.join_src)
      return seg2 ? launch_tc_split<32, 16, 42, 2, 0>(L, in, out, N, kind, st) : launch_tc_split<32, 16, 42, 1, 0>(L, in, out, N, kind, st);
    return cudaErrorInvalidValue;
  }
  if (L.amode == 2) return launch_tc<1, 16, 31, 1, 2, 8>(L, in, out, N, kind, st);
  if (cb == 1 && cp == 32 && pool == 41) return launch_tc<1, 32, 41, 1, 1>(L, in, out, N, kind, st);
  if (cb == 4 && cp == 32 && pool == 41)
    return L.join_src ? launch_tc<4, 32, 41, 1, 0, 32, true>(L, in, out, N, kind, st)
                      : launch_tc<4, 32, 41, 1, 0>(L, in, out, N, kind, st);
  if (cb == 4 && cp == 64 && pool == 42) return launch_tc<4, 64, 42, 1, 0>(L, in, out, N, kind, st);
  if (cb == 8 && cp == 64 && pool == 42)
    return L.join_src ? launch_tc<8, 64, 42, 1, 0, 64, true>(L, in, out, N, kind, st)
                      : launch_tc<8, 64, 42, 1, 0>(L, in, out, N, kind, st);
  if (cb == 8 && cp == 64 && pool == 0)
    return seg2 ? launch_tc<8, 64, 0, 2, 0>(L, in, out, N, kind, st) : launch_tc<8, 64, 0, 1, 0>(L, in, out, N, kind, st);
  if (cb == 16 && cp == 16 && pool == 42)
    return seg2 ? launch_tc<16, 16, 42, 2, 0>(L, in, out, N, kind, st) : launch_tc<16, 16, 42, 1, 0>(L, in, out, N, kind, st);
  return cudaErrorInvalidValue;
}

template <typename TIn>
cudaError_t Conv0PoolH(const TIn* in, const float* w, const float* b, void* out, int N, int S, HalfKind kind,
                       cudaStream_t st) {
  const int PS = S - 4;
  dim3 grid((PS + kC0W - 1) / kC0W, (PS + kC0H - 1) / kC0H, N);
  conv0_pool_kernel<TIn><<<grid, 256, 0, st>>>(in, w, b, static_cast<uint8_t*>(out), S, kind == HalfKind::kBF16);
  return cudaGetLastError();
}
template cudaError_t Conv0PoolH<float>(const float*, const float*, const float*, void*, int, int, HalfKind, cudaStream_t);
template cudaError_t Conv0PoolH<uint8_t>(const uint8_t*, const float*, const float*, void*, int, int, HalfKind,
                                         cudaStream_t);


bool TailFusedSupported(int s7, int channels) { return channels == kTailC && s7 <= kTailMaxSide && s7 >= 11; }

cudaError_t TailFused(const void* p7, int N, int S7, float in_scale, const float* w8, const float* b8, const float* w9,
                      const float* b9, const float* ja, const float* jb, const float* jc, const DenseParams& dp,
                      int flat_len, HalfKind kind, long long* top1, float* probs, float* logits, float* dbg8,
                      float* dbg9, cudaStream_t st, bool split) {
  TailParams tp{w8, b8, w9, b9, ja, jb, jc, dp, S7, in_scale, flat_len, split ? 1 : 0};
  const int S8 = (S7 - 2 - 4) / 2 + 1, S9 = (S8 - 2 - 4) / 2 + 1;
  size_t floats = 16 * S7 * S7 + 16 * (S7 - 2) * (S7 - 2) + 16 * S8 * S8 + 2 * 16 * S9 * S9 + 2 * 9 * 16 * 16;
  for (int l = 0; l < 4; ++l) floats += static_cast<size_t>(l == 0 ? flat_len : dp.out[l - 1]) * dp.out[l] + dp.out[l];
  const size_t bytes = floats * sizeof(float);
  cudaError_t e = cudaFuncSetAttribute(tail_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(bytes));
  if (e != cudaSuccess) return e;
  // small batches: one image per CTA is latency-critical -> 1024 threads; large batches: 256 threads, more CTAs per SM
  // one CTA per image: as many threads as keeps every CTA of the launch resident in one wave
  const int sms = SmCount();
  const int tail_threads = N <= sms ? 1024 : (N <= 2 * sms ? 512 : 256);
  e = LaunchPdl(tail_fused_kernel, dim3(N), dim3(tail_threads), bytes, st, N, static_cast<const uint16_t*>(p7), tp,
                static_cast<int>(kind == HalfKind::kBF16), top1, probs, logits, dbg8, dbg9);
  if (e != cudaSuccess) return e;
  return cudaGetLastError();
}

// Vector form of chunked_to_f32_kernel for fp16 tensors: one thread per (image row, chunk, pixel), pixels fastest, so
// the 16-byte chunk loads of a warp are one contiguous 512-byte run and every store fills whole 32-byte sectors.
__global__ void chunked_to_f32_vec_kernel(const uint4* __restrict__ in, float4* __restrict__ out, int N, int S, int C,
                                          float scale, int split) {
  const int CBl = C / 8, CBn = split ? 2 * CBl : CBl;
  const size_t total = static_cast<size_t>(N) * S * CBl * S;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int x = static_cast<int>(i % S);
    size_t r = i / S;
    const int cb = static_cast<int>(r % CBl);
    r /= CBl;  // r = n * S + y
    const size_t at = (r * CBn + cb) * S + x;
    const uint4 h = in[at];
    float v[8];
    const uint32_t hw[4] = {h.x, h.y, h.z, h.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&hw[q]));
      v[2 * q] = f.x;
      v[2 * q + 1] = f.y;
    }
    if (split) {
      const uint4 l = in[at + static_cast<size_t>(CBl) * S];
      const uint32_t lw[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&lw[q]));
        v[2 * q] += f.x;
        v[2 * q + 1] += f.y;
      }
    }
    float4* o = out + ((r * S + x) * C + cb * 8) / 4;
    o[0] = make_float4(v[0] * scale, v[1] * scale, v[2] * scale, v[3] * scale);
    o[1] = make_float4(v[4] * scale, v[5] * scale, v[6] * scale, v[7] * scale);
  }
}

// fp32 NHWC [N, S, S, C] * scale -> split chunked tensor with Cpad logical channels (hi planes, then lo planes, fp16)
__global__ void f32_to_split_kernel(const float* __restrict__ in, uint16_t* __restrict__ out, int N, int S, int C, int Cpad,
                                    float scale, int bf16) {
  const size_t total = static_cast<size_t>(N) * S * S * Cpad;
  const int CBl = Cpad / 8;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % Cpad);
    size_t r = i / Cpad;
    const int x = static_cast<int>(r % S);
    r /= S;
    const int y = static_cast<int>(r % S);
    const int n = static_cast<int>(r / S);
    const float v = c < C ? in[((static_cast<size_t>(n) * S + y) * S + x) * C + c] * scale : 0.f;
    const size_t at = ((((static_cast<size_t>(n) * S + y) * (2 * CBl) + c / 8) * S) + x) * 8 + (c % 8);
    if (bf16) {
      const __nv_bfloat16 h = __float2bfloat16_rn(v);
      const __nv_bfloat16 l = __float2bfloat16_rn(v - __bfloat162float(h));
      out[at] = *reinterpret_cast<const uint16_t*>(&h);
      out[at + static_cast<size_t>(CBl) * S * 8] = *reinterpret_cast<const uint16_t*>(&l);
    } else {
      const __half h = __float2half_rn(v);
      const __half l = __float2half_rn(v - __half2float(h));
      out[at] = *reinterpret_cast<const uint16_t*>(&h);
      out[at + static_cast<size_t>(CBl) * S * 8] = *reinterpret_cast<const uint16_t*>(&l);
    }
  }
}

cudaError_t F32ToSplitChunked(const float* in, void* out, int N, int S, int C, int Cpad, float scale, HalfKind kind,
                              cudaStream_t st) {
  size_t total = static_cast<size_t>(N) * S * S * Cpad;
  int blocks = static_cast<int>(std::min<size_t>((total + 255) / 256, static_cast<size_t>(SmCount()) * 16));
  f32_to_split_kernel<<<blocks, 256, 0, st>>>(in, static_cast<uint16_t*>(out), N, S, C, Cpad, scale,
                                              kind == HalfKind::kBF16 ? 1 : 0);
  return cudaGetLastError();
}

cudaError_t ChunkedToF32(const void* in, float* out, int N, int S, int Ch, HalfKind kind, float scale,
                         cudaStream_t st, bool split) {
  size_t total = static_cast<size_t>(N) * S * S * Ch;
  if (kind == HalfKind::kF16 && Ch % 8 == 0) {
    int vblocks = static_cast<int>(std::min<size_t>((total / 8 + 255) / 256, static_cast<size_t>(SmCount()) * 16));
    chunked_to_f32_vec_kernel<<<vblocks, 256, 0, st>>>(static_cast<const uint4*>(in), reinterpret_cast<float4*>(out), N, S, Ch,
                                                       scale, split ? 1 : 0);
    return cudaGetLastError();
  }
  int blocks = static_cast<int>(std::min<size_t>((total + 255) / 256, static_cast<size_t>(SmCount()) * 16));
  chunked_to_f32_kernel<<<blocks, 256, 0, st>>>(static_cast<const uint16_t*>(in), out, N, S, Ch,
                                                kind == HalfKind::kBF16, scale, split ? 1 : 0);
  return cudaGetLastError();
}

}  // namespace rn
