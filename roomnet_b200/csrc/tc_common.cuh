// Shared device helpers of the sm_100a tensor-core kernels (kernels_tc.cu, kernels_block2.cu): inline-PTX wrappers
// for mbarrier / TMA / tcgen05, the shared-memory + TMEM configuration of one conv layer and packed 16-bit math.
#pragma once
#include <cuda.h>  // CUtensorMap types only; the encoder is fetched through cudaGetDriverEntryPoint
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdlib>
#include <mutex>

namespace rn {
namespace {

constexpr int kTileM = 128;           // pixels per row tile (UMMA M)
constexpr int kPlanePx = 136;         // pixels per channel-chunk plane in a smem stage
constexpr int kLoadPx = 132;          // pixels fetched per plane row (128 + taps, 16B multiple)
constexpr int kSlackBytes = 128 * 1024;  // over-read slack behind every chunked tensor (row tails + one padding row)
constexpr int kSmemBudget = 227 * 1024;

// ------------------------------------------------------- work partition ----
// The persistent kernels see their work as one line of output rows: image group 0 rows 0..side-1, group 1 rows
// 0..side-1, ...  The CTAs form gangs of n_strips (CTA b: gang b / n_strips, column strip b % n_strips); gang g of G owns
// the contiguous range [g*T/G, (g+1)*T/G) of that line, cut into pieces at image boundaries, so every CTA gets the same
// number of rows whatever N is (dealing equal row blocks round-robin left 14 % of the SMs idle in the last round at 128
// images).  The CTAs of a gang walk the same rows at the same pace, so the input rows that neighbouring strips share
// are fetched from HBM once (a line over (image, strip) units was measured 14 % slower on conv2d_1: the second strip
// came 200 rows after the first and found its input rows evicted from L2).  A piece costs a few halo rows and a pipeline
// refill, so a range boundary that falls within `min_piece` rows of an image boundary is moved onto it.
__host__ __device__ inline int rn_snap_row(long long r, int side, int min_piece) {
  const int m = static_cast<int>(r % side);
  if (m < min_piece) return static_cast<int>(r - m);
  if (side - m < min_piece) return static_cast<int>(r + (side - m));
  return static_cast<int>(r);
}
__host__ __device__ inline void rn_gang_rows(int gang, int n_gangs, int total_rows, int side, int min_piece, int* lo,
                                             int* hi) {
  *lo = rn_snap_row(static_cast<long long>(gang) * total_rows / n_gangs, side, min_piece);
  *hi = rn_snap_row(static_cast<long long>(gang + 1) * total_rows / n_gangs, side, min_piece);
}
// Host side: how many gangs to use and the snapping distance for `total_rows` rows of `side`-row images.
inline void rn_plan_rows(int total_rows, int side, int n_strips, int max_ctas, int* n_gangs, int* min_piece) {
  int g = total_rows / 2;  // at least two output rows per CTA
  if (g > max_ctas / n_strips) g = max_ctas / n_strips;
  if (g < 1) g = 1;
  const int quota = total_rows / g;
  int mp = quota / 3;
  if (mp > 6) mp = 6;
  if (mp > side / 2) mp = side / 2;
  *n_gangs = g;
  *min_piece = mp;
}

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
// Same wait with a suspend-time hint: the warp sleeps in hardware until the phase completes (or the hint expires) instead
// of re-issuing try_wait; spinning warps otherwise take issue slots from the working warps of their scheduler.
__device__ __forceinline__ void mbar_wait_sleep(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(done)
        : "r"(bar), "r"(parity), "r"(20000u)
        : "memory");
  } while (!done);
}
// two fp32 values in one 64-bit register pair: add.f32x2 / fma.f32x2 (sm_100) take one issue slot for two lanes of math
typedef unsigned long long f32x2_t;
__device__ __forceinline__ f32x2_t f2_pack(float lo, float hi) {
  f32x2_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void f2_unpack(f32x2_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2_t f2_add(f32x2_t a, f32x2_t b) {
  f32x2_t r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2_t f2_mul(f32x2_t a, f32x2_t b) {
  f32x2_t r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2_t f2_fma(f32x2_t a, f32x2_t b, f32x2_t c) {
  f32x2_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tma_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
// 4-D tiled TMA load (cp.async.bulk.tensor): box -> shared memory, completion on an mbarrier
__device__ __forceinline__ void tma_tensor4_g2s(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, int c3,
                                                uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::
          "r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar)
      : "memory");
}
// Programmatic dependent launch: a kernel launched with the programmatic-stream-serialization attribute may start (and
// run its prologue: barrier init, TMEM allocation, weight loads) while its predecessor in the stream is still
// draining; pdl_wait() blocks until the predecessor grid has completed and its writes are visible, pdl_trigger() lets
// the successor grid start launching as SMs become free.  Both are no-ops for ordinary launches.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
// L2 prefetch of a 4-D box (no shared-memory destination, no barrier): issued a few row pairs ahead of the load itself
__device__ __forceinline__ void tma_prefetch4(const CUtensorMap* map, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
// 5-D variant: the extra dimension is the image, for tiles that hold two windows of two images
__device__ __forceinline__ void tma_tensor5_g2s(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, int c3,
                                                int c4, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], "
      "[%7];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

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
// POOL modes: 0 = none, 31 = 3x3/1, 41 = 4x4/1, 42 = 4x4/2.  The kernel stores the window SUM of
// saturate(conv/6): the factors k*k and 6 are folded into the consumer's weights by the host.
// AMODE: 0 = channel-chunk planes (Cin >= 16), 1 = Cin 8: pixel pairs form a K=16 step (LBO = 16 B),
//        2 = conv0: a 16-byte chunk holds pixels (x, x+1) x (c0,c1,c2,0); chunks x and x+2 (LBO = 32 B) form one
//            K=16 step that covers all three dx taps; the two k-steps are the hi and lo halves of the weights
// SPLIT (AMODE 0 only): the fp32-class path.  Activations are stored as hi + lo 16-bit pairs, planes
// [hi chunks 0..CB/2-1 | lo chunks 0..CB/2-1] (CB = PHYSICAL planes = 2 x logical chunks), weights as [Wh planes | Wl
// planes], and one input row takes three products into the same accumulators: xh*Wh, xl*Wh, xh*Wl (xl*Wl is below the
// fp32 noise floor).  Only the K-step enumeration differs from the plain layer: 4.5 x the logical chunk count steps.
template <int CB, int COUT, int AMODE, bool WINDOWS, bool SPLIT = false>
struct TcCfg {
  static_assert(!SPLIT || (AMODE == 0 && CB % 4 == 0), "split layers: hi and lo halves, each an even number of chunks");
  // windowed tiles are written by one tiled TMA box [CB][4 windows][32 px][8] -> dense 128-pixel planes
  static constexpr int kPlanePxT = WINDOWS ? 128 : kPlanePx;
  static constexpr int kPlaneBytesT = kPlanePxT * 16;
  static constexpr int kPlanes = AMODE == 0 ? 3 * CB : 4;       // weight planes
  static constexpr int kCBL = SPLIT ? CB / 2 : CB;               // logical channel chunks
  static constexpr int kKSteps = AMODE == 0 ? (SPLIT ? 9 * (kCBL / 2) : 3 * (CB / 2)) : 2;  // MMAs per input row
  static constexpr int kSlots = (512 / COUT) > 16 ? 16 : (512 / COUT);
  static constexpr int kLogSlots = kSlots == 16 ? 4 : 3;
  static_assert(kSlots == 16 || kSlots == 8, "ring size must be a power of two");
  static constexpr int kTmemCols = kSlots * COUT;
  static constexpr int kWBytes = kPlanes * 3 * COUT * 16;
  static constexpr int kRowBytes = CB * kPlaneBytesT;                       // one input row, all planes
  static constexpr int kStageBytes = 2 * kRowBytes + (WINDOWS ? 128 : 0);  // a stage holds a PAIR of input rows (+ tap over-read pad)
  static constexpr int kFixedBytes = kWBytes + 4 * COUT * 4 + 1024;  // bias + join A/B/C
  static constexpr int kStagesFit = (kSmemBudget - kFixedBytes) / kStageBytes;
  static constexpr int kStages = kStagesFit > 8 ? 8 : kStagesFit;
  static constexpr int kSmemBytes = kFixedBytes + kStages * kStageBytes;
  // (a split layer with 128 input channels - conv2d_7 - holds 128 KB of hi + lo planes per row pair: one stage, its loads
  // and MMAs alternate; the layer is 1.6 % of the FLOPs)
  static_assert(kStages >= (SPLIT ? 1 : 2), "not enough shared memory for a double-buffered input ring");
  // descriptor offsets (in 16-byte units) of k-step ks relative to the stage / weight base
  // split layers: step ks = product (ks / (3 * kCBL / 2)): 0 = xh*Wh, 1 = xl*Wh, 2 = xh*Wl; inside a product the order is
  // (dx, chunk pair) as in the plain layer
  static constexpr int kPairs = kCBL / 2 > 0 ? kCBL / 2 : 1;  // chunk pairs per tap
  __host__ __device__ static constexpr uint32_t a_off16(int ks) {
    if (AMODE == 1) return static_cast<uint32_t>(2 * ks);
    if (AMODE != 0) return 0u;
    const int prod = ks / (3 * kPairs), r = ks % (3 * kPairs);
    const int plane = 2 * (r % kPairs) + ((SPLIT && prod == 1) ? kCBL : 0);
    return static_cast<uint32_t>(plane * kPlanePxT + r / kPairs);
  }
  __host__ __device__ static constexpr uint32_t b_off16(int ks) {
    if (AMODE != 0) return static_cast<uint32_t>(2 * ks * 3 * COUT);
    const int prod = ks / (3 * kPairs), r = ks % (3 * kPairs);
    const int plane = (r / kPairs) * kCBL + 2 * (r % kPairs) + ((SPLIT && prod == 2) ? 3 * kCBL : 0);
    return static_cast<uint32_t>(plane * 3 * COUT);
  }
  static constexpr uint32_t kALbo16 = AMODE == 0 ? kPlanePxT : (AMODE == 1 ? 1 : 2);  // K-direction core-matrix stride / 16
  static constexpr uint32_t kBLbo16 = 3 * COUT;
};

template <bool BF16>
struct H2 {
  __device__ static __forceinline__ uint32_t pack(float a, float b) {
    if constexpr (BF16) {
      __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
      return *reinterpret_cast<uint32_t*>(&h);
    } else {
      __half2 h = __floats2half2_rn(a, b);
      return *reinterpret_cast<uint32_t*>(&h);
    }
  }
  __device__ static __forceinline__ float2 unpack(uint32_t v) {
    if constexpr (BF16) {
      return __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&v));
    } else {
      return __half22float2(*reinterpret_cast<__half2*>(&v));
    }
  }
  __device__ static __forceinline__ uint32_t splat(float x) { return pack(x, x); }
  __device__ static __forceinline__ uint32_t sub(uint32_t a, uint32_t b) {
    if constexpr (BF16) {
      __nv_bfloat162 r = __hsub2(*reinterpret_cast<__nv_bfloat162*>(&a), *reinterpret_cast<__nv_bfloat162*>(&b));
      return *reinterpret_cast<uint32_t*>(&r);
    } else {
      __half2 r = __hsub2(*reinterpret_cast<__half2*>(&a), *reinterpret_cast<__half2*>(&b));
      return *reinterpret_cast<uint32_t*>(&r);
    }
  }
  __device__ static __forceinline__ uint32_t fma(uint32_t a, uint32_t b, uint32_t c) {  // a*b + c, single rounding
    if constexpr (BF16) {
      __nv_bfloat162 r = __hfma2(*reinterpret_cast<__nv_bfloat162*>(&a), *reinterpret_cast<__nv_bfloat162*>(&b),
                                 *reinterpret_cast<__nv_bfloat162*>(&c));
      return *reinterpret_cast<uint32_t*>(&r);
    } else {
      __half2 r = __hfma2(*reinterpret_cast<__half2*>(&a), *reinterpret_cast<__half2*>(&b),
                          *reinterpret_cast<__half2*>(&c));
      return *reinterpret_cast<uint32_t*>(&r);
    }
  }
  __device__ static __forceinline__ uint32_t add(uint32_t a, uint32_t b) {
    if constexpr (BF16) {
      __nv_bfloat162 r = __hadd2(*reinterpret_cast<__nv_bfloat162*>(&a), *reinterpret_cast<__nv_bfloat162*>(&b));
      return *reinterpret_cast<uint32_t*>(&r);
    } else {
      __half2 r = __hadd2(*reinterpret_cast<__half2*>(&a), *reinterpret_cast<__half2*>(&b));
      return *reinterpret_cast<uint32_t*>(&r);
    }
  }
  // mixed-precision fma (sm_100: FHFMA): fp32 acc + (16-bit x) * (16-bit y) taken from the low / high half of a packed
  // pair - no unpack instructions, one full-rate fma-pipe slot per element
  __device__ static __forceinline__ float fhfma_lo(uint32_t x, uint32_t y, float acc) {
    float r;
    if constexpr (BF16) {
      asm("{.reg .b16 xl, xh, yl, yh; mov.b32 {xl, xh}, %1; mov.b32 {yl, yh}, %2; fma.rn.f32.bf16 %0, xl, yl, %3;}"
          : "=f"(r) : "r"(x), "r"(y), "f"(acc));
    } else {
      asm("{.reg .b16 xl, xh, yl, yh; mov.b32 {xl, xh}, %1; mov.b32 {yl, yh}, %2; fma.rn.f32.f16 %0, xl, yl, %3;}"
          : "=f"(r) : "r"(x), "r"(y), "f"(acc));
    }
    return r;
  }
  __device__ static __forceinline__ float fhfma_hi(uint32_t x, uint32_t y, float acc) {
    float r;
    if constexpr (BF16) {
      asm("{.reg .b16 xl, xh, yl, yh; mov.b32 {xl, xh}, %1; mov.b32 {yl, yh}, %2; fma.rn.f32.bf16 %0, xh, yh, %3;}"
          : "=f"(r) : "r"(x), "r"(y), "f"(acc));
    } else {
      asm("{.reg .b16 xl, xh, yl, yh; mov.b32 {xl, xh}, %1; mov.b32 {yl, yh}, %2; fma.rn.f32.f16 %0, xh, yh, %3;}"
          : "=f"(r) : "r"(x), "r"(y), "f"(acc));
    }
    return r;
  }
  // mixed-precision add (FHADD): fp32 acc + 16-bit x from the low / high half of a packed pair
  __device__ static __forceinline__ float fhadd_lo(uint32_t x, float acc) {
    float r;
    if constexpr (BF16) {
      asm("{.reg .b16 xl, xh; mov.b32 {xl, xh}, %1; add.rn.f32.bf16 %0, xl, %2;}" : "=f"(r) : "r"(x), "f"(acc));
    } else {
      asm("{.reg .b16 xl, xh; mov.b32 {xl, xh}, %1; add.rn.f32.f16 %0, xl, %2;}" : "=f"(r) : "r"(x), "f"(acc));
    }
    return r;
  }
  __device__ static __forceinline__ float fhadd_hi(uint32_t x, float acc) {
    float r;
    if constexpr (BF16) {
      asm("{.reg .b16 xl, xh; mov.b32 {xl, xh}, %1; add.rn.f32.bf16 %0, xh, %2;}" : "=f"(r) : "r"(x), "f"(acc));
    } else {
      asm("{.reg .b16 xl, xh; mov.b32 {xl, xh}, %1; add.rn.f32.f16 %0, xh, %2;}" : "=f"(r) : "r"(x), "f"(acc));
    }
    return r;
  }
};

// Warp-converged call: one elected lane issues D[tmem] += A[smem] * B[smem].  elect.sync (not `lane == 0`) lets ptxas
// emit ELECT + a predicated UTCHMMA instead of a per-active-lane serialisation loop around every MMA.
// Same elected lane (deterministic for a full mask) commits: arrive on `bar` once all its prior MMAs completed.
// variants for code that already runs on a single elected thread
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, q;\n\t"
      "}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tc_mma_acc1(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi,
                                            uint32_t idesc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .b64 da, db;\n\t"
      "setp.eq.b32 p, 0, 0;\n\t"
      "mov.b64 da, {%1, %3};\n\t"
      "mov.b64 db, {%2, %3};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t"
      "}" ::"r"(d_tmem),
      "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc)
      : "memory");
}
template <int N>
__device__ __forceinline__ void tc_ld(uint32_t taddr, float* v) {
  static_assert(N == 2 || N == 4 || N == 8 || N == 16 || N == 32, "unsupported TMEM load width");
  uint32_t r[N];
  if constexpr (N == 2) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(taddr));
  } else if constexpr (N == 4) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(taddr));
  } else if constexpr (N == 8) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
  } else if constexpr (N == 16) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
  } else {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
  }
#pragma unroll
  for (int i = 0; i < N; ++i) v[i] = __uint_as_float(r[i]);
}
template <int N>
__device__ __forceinline__ void tc_st(uint32_t taddr, const float* v) {
  static_assert(N == 2 || N == 4 || N == 8 || N == 16 || N == 32, "unsupported TMEM store width");
#define RN_U(i) "r"(__float_as_uint(v[i]))
  if constexpr (N == 2) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1,%2};" ::"r"(taddr), RN_U(0), RN_U(1) : "memory");
  } else if constexpr (N == 4) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr), RN_U(0), RN_U(1), RN_U(2),
                 RN_U(3)
                 : "memory");
  } else if constexpr (N == 8) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), RN_U(0),
                 RN_U(1), RN_U(2), RN_U(3), RN_U(4), RN_U(5), RN_U(6), RN_U(7)
                 : "memory");
  } else if constexpr (N == 16) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(
            taddr),
        RN_U(0), RN_U(1), RN_U(2), RN_U(3), RN_U(4), RN_U(5), RN_U(6), RN_U(7), RN_U(8), RN_U(9), RN_U(10), RN_U(11),
        RN_U(12), RN_U(13), RN_U(14), RN_U(15)
        : "memory");
  } else {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
        "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(taddr),
        RN_U(0), RN_U(1), RN_U(2), RN_U(3), RN_U(4), RN_U(5), RN_U(6), RN_U(7), RN_U(8), RN_U(9), RN_U(10), RN_U(11),
        RN_U(12), RN_U(13), RN_U(14), RN_U(15), RN_U(16), RN_U(17), RN_U(18), RN_U(19), RN_U(20), RN_U(21), RN_U(22),
        RN_U(23), RN_U(24), RN_U(25), RN_U(26), RN_U(27), RN_U(28), RN_U(29), RN_U(30), RN_U(31)
        : "memory");
  }
#undef RN_U
}
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// warp 0 TMA producer, warp 1 MMA issuer, then 4 epilogue warps (TMEM lane quadrants) per channel group
__host__ __device__ constexpr int tc_groups(int creal) { return creal >= 32 ? 4 : 2; }
__host__ __device__ constexpr int tc_threads(int creal) { return 64 + 128 * tc_groups(creal); }

// Launch with the programmatic-stream-serialization attribute (see pdl_wait): the kernel may be scheduled while its
// predecessor in the stream is finishing.  Used for small batches only (measured: batch-1 latency 0.167 -> 0.140 ms);
// with large micro-batches on two streams a pre-launched CTA would sit on an SM (shared memory, TMEM) that the other
// stream's kernels could have used (end-to-end throughput -3 %).
constexpr int kPdlMaxBatch = 32;
template <typename... KArgs, typename... Args>
cudaError_t LaunchPdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, int batch,
                      Args... args) {
  const bool enabled = batch <= kPdlMaxBatch;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = enabled ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, args...);
}

using PFN_encodeTiled = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                     const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                     CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// cuTensorMapEncodeTiled through the runtime's driver entry point table (no -lcuda); resolved once, thread-safe
// (launches are issued concurrently from one host thread per replica).
inline PFN_encodeTiled GetEncodeTiled() {
  static std::once_flag once;
  static PFN_encodeTiled fn_cached = nullptr;
  std::call_once(once, [] {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t ee = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (ee == cudaSuccess && qres == cudaDriverEntryPointSuccess && fn) fn_cached = reinterpret_cast<PFN_encodeTiled>(fn);
  });
  return fn_cached;
}

}  // namespace
}  // namespace rn
