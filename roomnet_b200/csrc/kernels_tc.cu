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
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <cstdio>
#include <cstring>
#include <vector>

#include "kernels.h"

namespace rn {

namespace {

constexpr int kTileM = 128;           // pixels per row tile (UMMA M)
constexpr int kPlanePx = 136;         // pixels per channel-chunk plane in a smem stage
constexpr int kPlaneBytes = kPlanePx * 16;
constexpr int kLoadPx = 132;          // pixels fetched per plane row (128 + taps, 16B multiple)
constexpr int kSlackBytes = 4096;     // over-read slack behind every chunked tensor
constexpr int kThreads = 192;         // warp 0 TMA producer, warp 1 MMA issuer, warps 2..5 epilogue
constexpr int kSmemBudget = 227 * 1024;

// ------------------------------------------------------------------ PTX ----
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tma_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// UMMA shared-memory descriptor, no swizzle, K-major:  ((8,m),(8,2)) : ((16B, SBO), (2B, LBO))
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4) | (static_cast<uint64_t>(lbo_bytes >> 4) << 16) |
         (static_cast<uint64_t>(sbo_bytes >> 4) << 32) | (1ull << 46);
}
// kind::f16 instruction descriptor: D=f32, A/B = f16 (0) or bf16 (1), both K-major, M=128.
__device__ __forceinline__ uint32_t make_idesc(int n, int bf16) {
  return (1u << 4) | (static_cast<uint32_t>(bf16) << 7) | (static_cast<uint32_t>(bf16) << 10) |
         (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(kTileM >> 4) << 24);
}

__device__ __forceinline__ float relu6f(float v) { return fminf(fmaxf(v, 0.f), 6.f); }

__device__ __forceinline__ uint32_t pack2(float a, float b, int bf16) {
  if (bf16) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  }
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float unpack_lo(uint32_t v, int bf16) {
  if (bf16) return __bfloat162float(reinterpret_cast<__nv_bfloat162*>(&v)->x);
  return __half2float(reinterpret_cast<__half2*>(&v)->x);
}
__device__ __forceinline__ float unpack_hi(uint32_t v, int bf16) {
  if (bf16) return __bfloat162float(reinterpret_cast<__nv_bfloat162*>(&v)->y);
  return __half2float(reinterpret_cast<__half2*>(&v)->y);
}

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
  int rows_per_item;  // pooled (output) rows per work item
  int n_rowblocks;
  int n_items;
  int w_bytes;        // packed weight bytes per part
  int bf16;
};

template <int CB, int COUT>
struct TcCfg {
  static constexpr bool kPaired = (CB == 1);                 // Cin = 8: two adjacent pixels form one K=16 step
  static constexpr int kPlanes = kPaired ? 4 : 3 * CB;       // weight planes (dx-major)
  static constexpr int kKSteps = kPaired ? 2 : 3 * (CB / 2); // MMAs per input row (before ring splits)
  static constexpr int kSlots = (512 / COUT) > 16 ? 16 : (512 / COUT);
  static constexpr int kTmemCols = kSlots * COUT;
  static constexpr int kWBytes = kPlanes * 3 * COUT * 16;
  static constexpr int kStageBytes = CB * kPlaneBytes;
  static constexpr int kFixedBytes = kWBytes + COUT * 4 + 2 * 4 * COUT * 16 + 1024;
  static constexpr int kStagesFit = (kSmemBudget - kFixedBytes) / kStageBytes;
  static constexpr int kStages = kStagesFit > 6 ? 6 : kStagesFit;
  static constexpr int kSmemBytes = kFixedBytes + kStages * kStageBytes;
  static_assert(kStages >= 3, "not enough shared memory for a 3-stage input ring");
};

struct Item {
  int n0, x_in0, x_out0, po0, npo, c0, nconv;
};

template <int POOL_S, int SEG>
__device__ __forceinline__ Item decode_item(const TcParams& p, int item) {
  Item it;
  int rb = item % p.n_rowblocks;
  int t = item / p.n_rowblocks;
  int strip = t % p.n_strips;
  int ig = t / p.n_strips;
  it.n0 = ig * SEG;
  it.x_in0 = strip * p.strip_step_in;
  it.x_out0 = strip * p.strip_step_out;
  it.po0 = rb * p.rows_per_item;
  it.npo = min(p.rows_per_item, p.out_side - it.po0);
  if (POOL_S == 0) {
    it.c0 = it.po0;
    it.nconv = it.npo;
  } else {
    it.c0 = it.po0 * POOL_S;
    it.nconv = (it.npo - 1) * POOL_S + 4;
  }
  return it;
}

// ---------------------------------------------------------------------------
// CB     : input channel chunks (Cin/8)           COUT : output channels of this pass
// POOL_S : 0 = no pooling, 1 = 4x4/1, 2 = 4x4/2   SEG  : images side by side in one 128-pixel tile
// ---------------------------------------------------------------------------
template <int CB, int COUT, int POOL_S, int SEG>
__global__ void __launch_bounds__(kThreads, 1) conv_tc_kernel(const TcParams p) {
  using Cfg = TcCfg<CB, COUT>;
  constexpr int R = Cfg::kSlots;
  constexpr int NST = Cfg::kStages;
  constexpr int SEGW = kTileM / SEG;
  constexpr uint32_t kStageTx = CB * (SEG == 1 ? kLoadPx * 16 : 2 * SEGW * 16);

  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* s_w = smem;                                            // packed weights
  uint8_t* s_stage = smem + Cfg::kWBytes;                         // NST input-row stages
  float* s_bias = reinterpret_cast<float*>(s_stage + NST * Cfg::kStageBytes);
  float4* s_xchg = reinterpret_cast<float4*>(s_bias + COUT);      // [2][4 quads][COUT] ghost columns
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_xchg + 2 * 4 * COUT);
  // barrier map: [0,NST) full, [NST,2NST) empty, 2NST weights, then R acc_full, R acc_empty
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_bar + 2 * NST + 1 + 2 * R);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar0 = smem_u32(s_bar);
  auto bar_full = [&](int s) { return bar0 + 8u * s; };
  auto bar_empty = [&](int s) { return bar0 + 8u * (NST + s); };
  const uint32_t bar_w = bar0 + 8u * (2 * NST);
  auto bar_accf = [&](int s) { return bar0 + 8u * (2 * NST + 1 + s); };
  auto bar_acce = [&](int s) { return bar0 + 8u * (2 * NST + 1 + R + s); };

  const int part = blockIdx.y;
  const uint8_t* w_gmem = p.w + static_cast<size_t>(part) * p.w_bytes;
  const float* bias_g = p.bias + part * COUT;

  if (threadIdx.x == 0) {
    for (int s = 0; s < NST; ++s) {
      mbar_init(bar_full(s), 1);
      mbar_init(bar_empty(s), 1);
    }
    mbar_init(bar_w, 1);
    for (int s = 0; s < R; ++s) {
      mbar_init(bar_accf(s), 1);
      mbar_init(bar_acce(s), 4);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)),
                 "r"(Cfg::kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = threadIdx.x; i < COUT; i += kThreads) s_bias[i] = bias_g[i];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;

  const size_t in_row_bytes = static_cast<size_t>(CB) * p.in_side * 16;  // all planes of one input row
  const size_t in_img_bytes = in_row_bytes * p.in_side;

  if (warp == 0) {
    // =========================== TMA producer ===========================
    if (lane == 0) {
      mbar_arrive_expect_tx(bar_w, Cfg::kWBytes);
      for (int off = 0; off < Cfg::kWBytes; off += 16384) {
        int sz = min(16384, Cfg::kWBytes - off);
        tma_bulk_g2s(smem_u32(s_w + off), w_gmem + off, sz, bar_w);
      }
      uint32_t cnt = 0;
      for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
        const Item it = decode_item<POOL_S, SEG>(p, item);
        const int nin = it.nconv + 2;
        for (int r = 0; r < nin; ++r, ++cnt) {
          const int st = cnt % NST;
          mbar_wait(bar_empty(st), ((cnt / NST) & 1) ^ 1);
          mbar_arrive_expect_tx(bar_full(st), kStageTx);
          const uint32_t dst = smem_u32(s_stage + st * Cfg::kStageBytes);
#pragma unroll
          for (int sg = 0; sg < SEG; ++sg) {
            const int n = min(it.n0 + sg, p.N - 1);
            const uint8_t* src = p.in + n * in_img_bytes + (it.c0 + r) * in_row_bytes +
                                 static_cast<size_t>(it.x_in0) * 16;
            for (int c = 0; c < CB; ++c)
              tma_bulk_g2s(dst + c * kPlaneBytes + sg * SEGW * 16, src + static_cast<size_t>(c) * p.in_side * 16,
                           SEG == 1 ? kLoadPx * 16 : SEGW * 16, bar_full(st));
          }
        }
      }
    }
  } else if (warp == 1) {
    // ============================ MMA issuer ============================
    mbar_wait(bar_w, 0);
    const uint32_t w_base = smem_u32(s_w);
    const uint32_t idesc0 = make_idesc(0, p.bf16);
    uint32_t cnt = 0, G = 0;
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
      const Item it = decode_item<POOL_S, SEG>(p, item);
      const int nin = it.nconv + 2;
      for (int r = 0; r < nin; ++r, ++cnt) {
        const int st = cnt % NST;
        if (r < it.nconv) {  // the accumulator of conv row r is (re)started by this input row
          const uint32_t gy = G + r;
          mbar_wait(bar_acce(gy % R), ((gy / R) & 1) ^ 1);
        }
        mbar_wait(bar_full(st), (cnt / NST) & 1);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t a_base = smem_u32(s_stage + st * Cfg::kStageBytes);
          const int jlo = max(0, 2 - r);                       // j = 2 - dy ; conv row y = r - 2 + j
          const int jhi = min(2, it.nconv + 1 - r);
#pragma unroll 1
          for (int ks = 0; ks < Cfg::kKSteps; ++ks) {
            uint64_t a_desc;
            int plane;
            if constexpr (Cfg::kPaired) {
              a_desc = make_desc(a_base + (2 * ks) * 16, 16, 128);
              plane = 2 * ks;
            } else {
              constexpr int kHalfCb = CB / 2;
              const int dx = ks / kHalfCb, kk = ks % kHalfCb;
              a_desc = make_desc(a_base + (2 * kk) * kPlaneBytes + dx * 16, kPlaneBytes, 128);
              plane = dx * CB + 2 * kk;
            }
            const bool first = (ks == 0);
            int j = jlo;
            while (j <= jhi) {
              const int s0 = (G + r - 2 + j) % R;
              const bool fresh = (j == 2) && first;
              int len = 1;
              if (!fresh)
                while (j + len <= jhi && s0 + len < R && !((j + len) == 2 && first)) ++len;
              const uint64_t b_desc =
                  make_desc(w_base + (plane * 3 * COUT + j * COUT) * 16, 3 * COUT * 16, 128);
              tc_mma_f16(tmem_base + s0 * COUT, a_desc, b_desc, idesc0 | (static_cast<uint32_t>((len * COUT) >> 3) << 17),
                         fresh ? 0u : 1u);
              j += len;
            }
          }
          tc_commit(bar_empty(st));
          if (r >= 2) tc_commit(bar_accf((G + r - 2) % R));
        }
        __syncwarp();
      }
      G += it.nconv;
    }
  } else {
    // ============================= epilogue =============================
    const int quad = warp & 3;                    // TMEM lane quadrant this warp may read
    const int pix = quad * 32 + lane;             // pixel (UMMA row) owned by this thread
    const int seg = pix / SEGW, xs = pix % SEGW;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    const int bf16 = p.bf16;
    const size_t out_row_bytes = static_cast<size_t>(p.cb_out_total) * p.out_side * 16;
    const size_t out_img_bytes = out_row_bytes * p.out_side;
    uint32_t G = 0;
    uint32_t xrow = 0;  // output rows emitted (ghost buffer parity)

    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
      const Item it = decode_item<POOL_S, SEG>(p, item);
      const int n_img = it.n0 + seg;
      bool col_ok;
      int col;
      if (POOL_S == 0) {
        col = it.x_in0 + xs;
        col_ok = col < p.conv_side;
      } else if (POOL_S == 1) {
        col = it.x_out0 + xs;
        col_ok = (xs + 3 < SEGW) && (xs < p.strip_step_out || p.n_strips == 1) && col < p.out_side;
      } else {
        col = it.x_out0 + (xs >> 1);
        col_ok = !(xs & 1) && (xs + 3 < SEGW) && ((xs >> 1) < p.strip_step_out || p.n_strips == 1) && col < p.out_side;
      }
      col_ok = col_ok && n_img < p.N;
      uint8_t* out_px = p.out + n_img * out_img_bytes + (static_cast<size_t>(part) * (COUT / 8) * p.out_side + col) * 16;

      // vertical pooling window (registers).  POOL_S==1: r1 = previous row, q1/q2 = pair sums ending 1/2 rows back.
      float r1[POOL_S == 1 ? COUT : 1], q1[POOL_S != 0 ? COUT : 1], q2[POOL_S == 1 ? COUT : 1];
#pragma unroll
      for (int c = 0; c < (POOL_S != 0 ? COUT : 1); ++c) q1[c] = 0.f;
#pragma unroll
      for (int c = 0; c < (POOL_S == 1 ? COUT : 1); ++c) r1[c] = q2[c] = 0.f;

      constexpr int kRowStep = POOL_S == 2 ? 2 : 1;
      for (int y = 0; y < it.nconv; y += kRowStep) {
        const uint32_t gy = G + y;
        const int slot0 = gy % R;
        mbar_wait(bar_accf(slot0), (gy / R) & 1);
        int slot1 = 0;
        if (POOL_S == 2) {
          slot1 = (gy + 1) % R;
          mbar_wait(bar_accf(slot1), ((gy + 1) / R) & 1);
        }
        tc_fence_after();
        bool emit;
        int yo;
        if (POOL_S == 0) {
          emit = true;
          yo = it.c0 + y;
        } else if (POOL_S == 1) {
          emit = y >= 3;
          yo = it.po0 + y - 3;
        } else {
          emit = y >= 2;
          yo = it.po0 + (y - 2) / 2;
        }
        float4* xw = s_xchg + ((xrow & 1) * 4 + quad) * COUT;
        const float4* xr = s_xchg + ((xrow & 1) * 4 + ((quad + 1) & 3)) * COUT;

        // pass 1: TMEM -> registers, bias + ReLU6, vertical window; keeps the vertical sums in v[]
        float v[COUT];
#pragma unroll
        for (int g = 0; g < COUT / 16; ++g) {
          uint32_t a[16], b[16];
          tc_ld16(t_lane + slot0 * COUT + g * 16, a);
          if (POOL_S == 2) tc_ld16(t_lane + slot1 * COUT + g * 16, b);
          tc_wait_ld();
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            const int c = g * 16 + e;
            const float bia = s_bias[c];
            float x = relu6f(__uint_as_float(a[e]) + bia);
            if (POOL_S == 0) {
              v[c] = x;
            } else if (POOL_S == 1) {
              const float q0 = x + r1[c];
              v[c] = q0 + q2[c];
              q2[c] = q1[c];
              q1[c] = q0;
              r1[c] = x;
            } else {
              const float q0 = x + relu6f(__uint_as_float(b[e]) + bia);
              v[c] = q0 + q1[c];
              q1[c] = q0;
            }
          }
        }
        // accumulator slots are free again
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(bar_acce(slot0));
          if (POOL_S == 2) mbar_arrive(bar_acce(slot1));
        }
        if (!emit) continue;
        if (POOL_S != 0) {
          // ghost columns: the first three pixels of the next quadrant
          if (lane < 3) {
#pragma unroll
            for (int c = 0; c < COUT; ++c) reinterpret_cast<float*>(&xw[c])[lane] = v[c];
          }
          asm volatile("bar.sync 1, 128;" ::: "memory");
        }
        uint8_t* orow = out_px + yo * out_row_bytes;
#pragma unroll
        for (int cb = 0; cb < COUT / 8; ++cb) {
          float h[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const int c = cb * 8 + e;
            if (POOL_S == 0) {
              h[e] = v[c];
            } else {
              const float4 gh = xr[c];
              float a1 = __shfl_down_sync(0xffffffffu, v[c], 1);
              if (lane == 31) a1 = gh.x;
              const float t = v[c] + a1;
              float b2 = __shfl_down_sync(0xffffffffu, t, 2);
              if (lane == 30) b2 = gh.x + gh.y;
              if (lane == 31) b2 = gh.y + gh.z;
              h[e] = (t + b2) * 0.0625f;
            }
          }
          if (col_ok) {
            uint4 o;
            o.x = pack2(h[0], h[1], bf16);
            o.y = pack2(h[2], h[3], bf16);
            o.z = pack2(h[4], h[5], bf16);
            o.w = pack2(h[6], h[7], bf16);
            *reinterpret_cast<uint4*>(orow + static_cast<size_t>(cb) * p.out_side * 16) = o;
          }
        }
        ++xrow;
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
      h[o] = s / 9.f;
    }
    uint4 o4;
    o4.x = pack2(h[0], h[1], bf16);
    o4.y = pack2(h[2], h[3], bf16);
    o4.z = pack2(h[4], h[5], bf16);
    o4.w = pack2(h[6], h[7], bf16);
    *reinterpret_cast<uint4*>(out + ((static_cast<size_t>(n) * PS + oy) * PS + ox) * 16) = o4;
  }
}

// out = A*p + B*resize_bilinear_legacy(src) + C on chunked tensors; one thread = one 8-channel chunk.
__global__ void join_h_kernel(const uint4* __restrict__ p, const uint4* __restrict__ src, uint4* __restrict__ out,
                              const float* __restrict__ A, const float* __restrict__ B, const float* __restrict__ Cc,
                              int N, int S, int SS, int CBn, int bf16) {
  const size_t total = static_cast<size_t>(N) * S * CBn * S;
  const float scale = static_cast<float>(SS) / static_cast<float>(S);
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    int x = static_cast<int>(i % S);
    size_t r = i / S;
    int cb = static_cast<int>(r % CBn);
    r /= CBn;
    int y = static_cast<int>(r % S);
    int n = static_cast<int>(r / S);
    float fy = static_cast<float>(y) * scale, fx = static_cast<float>(x) * scale;
    int y0 = static_cast<int>(floorf(fy)), x0 = static_cast<int>(floorf(fx));
    int y1 = min(y0 + 1, SS - 1), x1 = min(x0 + 1, SS - 1);
    float ty = fy - static_cast<float>(y0), tx = fx - static_cast<float>(x0);
    const uint4* s_n = src + static_cast<size_t>(n) * SS * CBn * SS;
    auto at = [&](int yy, int xx) { return s_n[(static_cast<size_t>(yy) * CBn + cb) * SS + xx]; };
    const uint4 tl = at(y0, x0), tr = at(y0, x1), bl = at(y1, x0), br = at(y1, x1);
    const uint4 pv = p[i];
    const uint32_t* ptl = &tl.x;
    const uint32_t* ptr = &tr.x;
    const uint32_t* pbl = &bl.x;
    const uint32_t* pbr = &br.x;
    const uint32_t* ppv = &pv.x;
    uint4 o;
    uint32_t* po = &o.x;
#pragma unroll
    for (int e2 = 0; e2 < 4; ++e2) {
      float res[2];
#pragma unroll
      for (int hl = 0; hl < 2; ++hl) {
        auto get = [&](const uint32_t* q) { return hl ? unpack_hi(q[e2], bf16) : unpack_lo(q[e2], bf16); };
        const int c = cb * 8 + e2 * 2 + hl;
        float a = get(ptl), b = get(ptr), cc = get(pbl), d = get(pbr);
        float top = a + (b - a) * tx;
        float bot = cc + (d - cc) * tx;
        float rs = top + (bot - top) * ty;
        res[hl] = fmaf(A[c], get(ppv), fmaf(B[c], rs, Cc[c]));
      }
      po[e2] = pack2(res[0], res[1], bf16);
    }
    out[i] = o;
  }
}

__global__ void chunked_to_f32_kernel(const uint16_t* __restrict__ in, float* __restrict__ out, int N, int S, int C,
                                      int bf16) {
  const size_t total = static_cast<size_t>(N) * S * S * C;
  const int CBn = C / 8;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    int c = static_cast<int>(i % C);
    size_t r = i / C;
    int x = static_cast<int>(r % S);
    r /= S;
    int y = static_cast<int>(r % S);
    int n = static_cast<int>(r / S);
    uint16_t raw = in[((((static_cast<size_t>(n) * S + y) * CBn + c / 8) * S) + x) * 8 + (c % 8)];
    float f;
    if (bf16) {
      f = __uint_as_float(static_cast<uint32_t>(raw) << 16);
    } else {
      __half h = *reinterpret_cast<__half*>(&raw);
      f = __half2float(h);
    }
    out[i] = f;
  }
}

template <int CB, int COUT, int POOL_S, int SEG>
cudaError_t launch_tc(const TcConvLayer& L, const void* in, void* out, int N, HalfKind kind, cudaStream_t st) {
  using Cfg = TcCfg<CB, COUT>;
  TcParams p{};
  p.in = static_cast<const uint8_t*>(in);
  p.out = static_cast<uint8_t*>(out);
  p.w = static_cast<const uint8_t*>(L.w_packed);
  p.bias = L.bias;
  p.N = N;
  p.in_side = L.in_side;
  p.conv_side = L.in_side - 2;
  p.out_side = L.out_side;
  p.cb_out_total = L.cout / 8;
  p.w_bytes = static_cast<int>(L.w_bytes);
  p.bf16 = kind == HalfKind::kBF16;
  constexpr int SEGW = kTileM / SEG;
  if (SEG == 2 && L.in_side > SEGW) return cudaErrorInvalidValue;
  if (POOL_S == 0) {
    p.strip_step_in = p.strip_step_out = kTileM;
    p.n_strips = SEG == 2 ? 1 : (p.conv_side + kTileM - 1) / kTileM;
  } else if (POOL_S == 1) {
    p.strip_step_in = p.strip_step_out = kTileM - 3;
    p.n_strips = SEG == 2 ? 1 : (p.out_side + p.strip_step_out - 1) / p.strip_step_out;
  } else {
    p.strip_step_out = (kTileM - 4) / 2 + 1;  // 63 pooled columns per strip
    p.strip_step_in = 2 * p.strip_step_out;
    p.n_strips = SEG == 2 ? 1 : (p.out_side + p.strip_step_out - 1) / p.strip_step_out;
  }
  const int groups = (N + SEG - 1) / SEG;
  const int base_items = groups * p.n_strips;
  // split rows so that the persistent grid sees >= ~4 waves, but keep >= 16 pooled rows per item
  int nrb = (4 * 148 + base_items - 1) / base_items;
  nrb = std::max(1, std::min(nrb, std::max(1, p.out_side / 16)));
  p.rows_per_item = (p.out_side + nrb - 1) / nrb;
  p.n_rowblocks = (p.out_side + p.rows_per_item - 1) / p.rows_per_item;
  p.n_items = base_items * p.n_rowblocks;
  auto kern = conv_tc_kernel<CB, COUT, POOL_S, SEG>;
  // per device (replicas of several GPUs share the process), and cheap enough to repeat
  cudaError_t ea = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
  if (ea != cudaSuccess) return ea;
  dim3 grid(std::min(p.n_items, 148), L.cout_parts);
  kern<<<grid, kThreads, Cfg::kSmemBytes, st>>>(p);
  return cudaGetLastError();
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

size_t PackTcWeights(const double* w, int cin, int cout, int cout_parts, HalfKind kind, void* out_host) {
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
            double v = w[(((dy * 3 + dx) * cin) + c0 + e) * cout + part * cp + oc];
            o[part * (part_bytes / 2) + ((static_cast<size_t>(pl) * 3 * cp + j * cp + oc) * 8) + e] =
                to_half_bits(v, kind);
          }
      }
    }
  return part_bytes;
}

cudaError_t ConvTc(const TcConvLayer& L, const void* in, void* out, int N, HalfKind kind, cudaStream_t st) {
  const int cb = L.cin / 8, cp = L.cout / L.cout_parts;
  const bool seg2 = L.in_side <= 64;
  if (cb == 1 && cp == 32 && L.pool_s == 1) return launch_tc<1, 32, 1, 1>(L, in, out, N, kind, st);
  if (cb == 4 && cp == 32 && L.pool_s == 1) return launch_tc<4, 32, 1, 1>(L, in, out, N, kind, st);
  if (cb == 4 && cp == 64 && L.pool_s == 2) return launch_tc<4, 64, 2, 1>(L, in, out, N, kind, st);
  if (cb == 8 && cp == 64 && L.pool_s == 2) return launch_tc<8, 64, 2, 1>(L, in, out, N, kind, st);
  if (cb == 8 && cp == 64 && L.pool_k == 0)
    return seg2 ? launch_tc<8, 64, 0, 2>(L, in, out, N, kind, st) : launch_tc<8, 64, 0, 1>(L, in, out, N, kind, st);
  if (cb == 16 && cp == 16 && L.pool_s == 2)
    return seg2 ? launch_tc<16, 16, 2, 2>(L, in, out, N, kind, st) : launch_tc<16, 16, 2, 1>(L, in, out, N, kind, st);
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

cudaError_t JoinH(const void* p, const void* src, void* out, const float* A, const float* B, const float* C, int N,
                  int S, int SS, int Ch, HalfKind kind, cudaStream_t st) {
  size_t total = static_cast<size_t>(N) * S * S * (Ch / 8);
  int blocks = static_cast<int>(std::min<size_t>((total + 255) / 256, 148 * 16));
  join_h_kernel<<<blocks, 256, 0, st>>>(static_cast<const uint4*>(p), static_cast<const uint4*>(src),
                                        static_cast<uint4*>(out), A, B, C, N, S, SS, Ch / 8, kind == HalfKind::kBF16);
  return cudaGetLastError();
}

cudaError_t ChunkedToF32(const void* in, float* out, int N, int S, int Ch, HalfKind kind, cudaStream_t st) {
  size_t total = static_cast<size_t>(N) * S * S * Ch;
  int blocks = static_cast<int>(std::min<size_t>((total + 255) / 256, 148 * 16));
  chunked_to_f32_kernel<<<blocks, 256, 0, st>>>(static_cast<const uint16_t*>(in), out, N, S, Ch,
                                                kind == HalfKind::kBF16);
  return cudaGetLastError();
}

}  // namespace rn
