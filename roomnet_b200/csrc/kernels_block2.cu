// Residual block 2 of RoomNet as ONE kernel for sm_100a (reference network.py:183-203, instantiated :227):
//
//     R2 --conv2d_2 3x3 + ReLU6 + avgpool 4/1--> P2 --conv2d_3 3x3 + ReLU6 + avgpool 4/1--> P3
//     J  = A*P3 + B*resize_bilinear(R2 -> size of P3) + C            (BN folded)
//
// R2 (the block's first pooled tensor, written by the conv2d_1 kernel) is read from HBM once, P2 never leaves the
// SM: the first layer's epilogue writes it straight into the shared-memory A-operand tiles of the second layer.
// Only J is written back.  Both layers are "row-stationary tap-stacked" tcgen05 implicit GEMMs exactly as in
// kernels_tc.cu (M = 128 pixels of one input row as four 32-pixel windows, N = 3 x 32 stacked dy taps, fp32
// accumulators in a TMEM ring of 8 row slots per layer: 2 x 8 x 32 = all 512 TMEM columns).
//
// Geometry of one work item = (image, column strip, block of output rows):
//   * a strip owns 103 output columns.  Layer 1 reads R2 through a tiled TMA box of four overlapping windows
//     (stride 27 pixels) and yields 4 x 27 = 108 P2 columns; its epilogue writes every P2 pixel into the one or two
//     lanes of the layer-2 tile that need it (window offsets {0, 27, 54, 76}: the last window overlaps its
//     neighbour a little more, so that 108 input columns give the maximal 103 output columns and 205/210/215-wide
//     rows of the 224x224 network are exactly two strips per row for every layer - no halo recomputation).
//   * rows stream top to bottom; per row pair the kernel runs: TMA (R2 pair) -> 12 MMAs -> epilogue 1 (clip,
//     4x4 window sum, 16-bit pack, st.shared into a P2 stage) -> 12 MMAs -> epilogue 2 (clip, window sum, residual
//     join, global store).
//   * the residual source rows are staged by TMA bulk copies into a small shared-memory ring (8 rows x 112 pixels)
//     a few row pairs ahead of their use, so the join costs shared-memory loads, not L2 gathers.
//
// Warp roles (one thread each): warp 0 = TMA producer of the R2 row pairs, warp 1 = MMA issuer of layer 1,
// warp 2 = TMA producer of the residual rows, warp 3 = MMA issuer of layer 2 - four independent dataflow loops
// coupled only through mbarriers; this first warp group hands most of its registers to the others (setmaxnreg).
// Warps 4..19 = epilogue: warp (quadrant q, channel group g) owns TMEM lanes 32q..32q+31 and
// channels 8g..8g+7 of BOTH layers and alternates between them (layer 2 runs kLagEpi row pairs behind layer 1).
#include <cstring>
#include <type_traits>

#include "kernels.h"
#include "tc_common.cuh"

namespace rn {

namespace {

using B2Cfg = TcCfg<4, 32, 0, true>;  // Cin = 32 (4 chunks), Cout = 32, windowed 128-pixel planes
constexpr int kB2R = 8;               // accumulator row slots per layer
constexpr int kB2RP = kB2R / 2;       // ... handled as pairs
constexpr int kB2LogR = 3;
constexpr int kB2NS1 = 4;             // R2 row-pair stages (TMA -> layer-1 MMA)
constexpr int kB2NS2 = 4;             // P2 row-pair stages (epilogue 1 -> layer-2 MMA)
constexpr int kB2ResRows = 8;         // residual ring: rows
constexpr int kB2ResPx = 112;         // ... pixels per row (103 output columns * scale + taps)
constexpr int kB2ResPlaneBytes = kB2ResPx * 16;
constexpr int kB2ResRowBytes = 4 * kB2ResPlaneBytes;
constexpr int kB2ResGroups = 4;       // barrier ring of the residual stream (one group = rows of one epilogue-2 step)
constexpr int kB2StripOut = 103;      // output columns owned by a strip
constexpr int kB2MaxStrips = 8;
constexpr int kB2LagEpi = 5;          // epilogue warps: layer-2 pair m is drained after layer-1 pair m + 5
constexpr int kB2Threads = 128 + 16 * 32;  // warp group 0: two TMA producers + two MMA issuers; 16 epilogue warps
// Register budget: 20 warps = 5 per scheduler, 16384 registers per scheduler -> 96 per thread at launch.  The first
// warp group gives most of its registers back (setmaxnreg.dec), the epilogue warp groups take them (setmaxnreg.inc).
constexpr int kB2RegsCtl = 32, kB2RegsEpi = 112;
static_assert(4 * kB2RegsCtl + 16 * kB2RegsEpi <= 20 * 96, "setmaxnreg can only redistribute the registers the CTA was launched with");
constexpr int kB2Bars = 2 * kB2NS1 + 2 * kB2NS2 + 1 + 4 * kB2RP + 2 * kB2ResGroups;
constexpr int kB2SmemBytes = 2 * B2Cfg::kWBytes + (kB2NS1 + kB2NS2) * B2Cfg::kStageBytes + kB2ResRows * kB2ResRowBytes +
                             (2 * 32 + 3 * 32) * 4 + kB2ResGroups * 16 + kB2Bars * 8 + 16;
static_assert(kB2SmemBytes <= kSmemBudget, "block-2 kernel does not fit in shared memory");
static_assert((kB2NS1 & (kB2NS1 - 1)) == 0 && (kB2NS2 & (kB2NS2 - 1)) == 0, "stage rings are indexed with masks");

struct B2Params {
  const uint8_t* in;   // R2, chunked [n][y][4][x][8], side in_side
  uint8_t* out;        // J,  chunked [n][y][4][x][8], side out_side
  const uint8_t* w1;   // packed weights of the two layers (PackTcWeights)
  const uint8_t* w2;
  const float* bias1;  // [32] each, already divided by 6
  const float* bias2;
  const float* abc;    // [3][32] join coefficients A, B, C
  int N, in_side, out_side;
  int n_strips;
  int total_rows, min_piece;  // work line of N x out_side output rows (tc_common.cuh: rn_gang_rows)
  float res_scale;     // in_side / out_side as float32 (TF computes the resize scale in float32)
  int x0[kB2MaxStrips];      // first column a strip computes (input, P2 and output columns share the origin)
  int own_lo[kB2MaxStrips];  // output columns [own_lo, own_hi) are stored by the strip
  int own_hi[kB2MaxStrips];
};

struct B2Maps {
  CUtensorMap m[kB2MaxStrips];  // per strip: R2 windows {256 el | 4 windows, stride 27 px | 4 planes | N*in_side rows}
  CUtensorMap res;              // R2 as 8-byte elements {2 * in_side | 4 planes | N*in_side rows}: box = 112 px x 4 planes
};

struct B2Item {
  int n, strip, po0, npo;
  int nconv3, nin2, nconv2, nin1;  // conv rows / input rows of the two layers (all even)
  int n1s, n1e, n2s, n2e;          // R2 stage loads, epilogue-1 steps, P2 pairs, epilogue-2 steps
};

// the piece of the CTA's row range [.., hi) that starts at line position `cur`
__device__ __forceinline__ B2Item b2_decode(const B2Params& p, int cur, int hi) {
  B2Item it;
  it.n = cur / p.out_side;
  it.strip = blockIdx.x % p.n_strips;
  it.po0 = cur - it.n * p.out_side;
  it.npo = min(p.out_side - it.po0, hi - cur);
  it.nconv3 = (it.npo + 3 + 1) & ~1;  // conv rows of layer 2 (one never-stored extra row when odd)
  it.nin2 = it.nconv3 + 2;            // P2 rows layer 2 reads = pooled rows layer 1 must produce
  it.nconv2 = it.nin2 + 4;            // conv rows of layer 1 (nin2 + 3, rounded up to even)
  it.nin1 = it.nconv2 + 2;            // R2 rows layer 1 reads
  it.n1s = it.nin1 >> 1;
  it.n1e = it.nconv2 >> 1;
  it.n2s = it.nin2 >> 1;
  it.n2e = it.nconv3 >> 1;
  return it;
}

__device__ __forceinline__ void tma_tensor3_g2s(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::
          "r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
      : "memory");
}

// One row PAIR of a tap-stacked layer: D[rows r0-2 .. r0+1] += A[input rows r0, r0+1] x [W(dy=2)|W(dy=1)|W(dy=0)].
// Same issue logic as conv_tc_kernel (kernels_tc.cu): interior pairs whose accumulator slots do not wrap around the
// ring are 12 MMAs with compile-time descriptor offsets, everything else goes through the general path.
__device__ __forceinline__ void b2_mma_pair(uint32_t a_lo, uint32_t b_lo0, uint32_t tmem_l, uint32_t idesc0, uint32_t G,
                                            int r0, int nconv) {
  using Cfg = B2Cfg;
  constexpr int R = kB2R, COUT = 32;
  constexpr uint32_t kDescHi = (128u >> 4) | (1u << 14);  // SBO = 128 B, descriptor version 1, SWIZZLE_NONE
  const uint32_t sb0 = (G + r0 - 2) & (R - 1);
  if (r0 >= 2 && r0 + 2 <= nconv && sb0 + 4 <= static_cast<uint32_t>(R)) {
    const uint32_t d0 = tmem_l + sb0 * COUT;
    constexpr uint32_t kIdescFull = static_cast<uint32_t>((3 * COUT) >> 3) << 17;
#pragma unroll
    for (int ks = 0; ks < Cfg::kKSteps; ++ks)
      tc_mma_acc1(d0, a_lo + Cfg::a_off16(ks), b_lo0 + Cfg::b_off16(ks), kDescHi, idesc0 | kIdescFull);
#pragma unroll
    for (int ks = 0; ks < Cfg::kKSteps; ++ks)
      tc_mma_acc1(d0 + COUT, a_lo + (Cfg::kRowBytes >> 4) + Cfg::a_off16(ks), b_lo0 + Cfg::b_off16(ks), kDescHi,
                  idesc0 | kIdescFull);
  } else {
#pragma unroll
    for (int sub = 0; sub < 2; ++sub) {
      const int r = r0 + sub;
      const int jlo = max(0, 2 - r);  // j = 2 - dy ; conv row y = r - 2 + j
      const int jhi = min(2, nconv + 1 - r);
      const uint32_t sb = (G + r - 2 + jlo) & (R - 1);
      const int nj = jhi - jlo + 1;
      const int len1 = min(nj, R - static_cast<int>(sb));  // slots before the ring wraps
      const uint32_t a_row = a_lo + sub * (Cfg::kRowBytes >> 4);
      {
        const uint32_t d = tmem_l + sb * COUT;
        const uint32_t idesc = idesc0 | (static_cast<uint32_t>((len1 * COUT) >> 3) << 17);
        const uint32_t b_lo = b_lo0 + jlo * COUT;
#pragma unroll
        for (int ks = 0; ks < Cfg::kKSteps; ++ks)
          tc_mma_acc1(d, a_row + Cfg::a_off16(ks), b_lo + Cfg::b_off16(ks), kDescHi, idesc);
      }
      if (len1 < nj) {  // ring wrap: the remaining conv rows start again at slot 0
        const uint32_t idesc = idesc0 | (static_cast<uint32_t>(((nj - len1) * COUT) >> 3) << 17);
        const uint32_t b_lo = b_lo0 + (jlo + len1) * COUT;
#pragma unroll
        for (int ks = 0; ks < Cfg::kKSteps; ++ks)
          tc_mma_acc1(tmem_l, a_row + Cfg::a_off16(ks), b_lo + Cfg::b_off16(ks), kDescHi, idesc);
      }
    }
  }
}

// saturate as a volatile statement: keeps ptxas from hoisting the clip of the second row above the last use of the
// state register it is written into (it would then need a copy to get the value there)
__device__ __forceinline__ float sat_here(float v) {
  float r;
  asm volatile("add.sat.f32 %0, %1, 0f00000000;" : "=f"(r) : "f"(v));
  return r;
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}

// in-place packed add (tied operands: the sum replaces `acc` in its own register pair)
__device__ __forceinline__ void f2_add_to(f32x2_t& acc, f32x2_t b) { asm("add.rn.f32x2 %0, %0, %1;" : "+l"(acc) : "l"(b)); }

// Drain one accumulator row pair (8 channels of this thread's pixel), hand the slots back pre-loaded with the bias,
// clip (saturate = ReLU6/6, the factor lives in the weights), advance the vertical 4-row window and apply the
// horizontal 4-column window on packed 16-bit pairs with warp shuffles.
// hp[0] = pooled row (y - 3), hp[1] = pooled row (y - 2) for conv rows (y, y + 1) of this call.
// The vertical window is a sum of two row-pair sums in packed fp32 (add.f32x2): with p(k) = x(k) + x(k+1),
// pooled(y-3) = p(y-3) + p(y-1) and pooled(y-2) = p(y-2) + p(y): four adds per channel pair for two output rows.
// State on entry: U = x(y-1), P[PH] = p(y-3), Q[PH] = p(y-2).  The call writes the new pair sums into P[PH^1] / Q[PH^1]
// and finishes the window sums in place in P[PH] / Q[PH], so the next call runs with the opposite phase and no
// register is ever copied (ptxas does not find this rotation by itself: ~18 moves per step otherwise).
// SCALE: the window sums are multiplied by the fp32 per-channel factors at shared address `scale_addr` before they are
// rounded to 16 bits (the join coefficient A of the block's last layer: exact, and the horizontal sums then run on A*x).
template <typename HH, int PH, bool SCALE>
__device__ __forceinline__ void b2_drain_pool(uint32_t t_slot0, uint32_t bias_addr, uint32_t bar_free, bool lane0,
                                              f32x2_t (&U)[4], f32x2_t (&P)[2][4], f32x2_t (&Q)[2][4],
                                              uint32_t (&hp)[2][4], uint32_t scale_addr = 0) {
  float a[8], b[8];
  tc_ld<8>(t_slot0, a);
  tc_ld<8>(t_slot0 + 32, b);
  tc_wait_ld();
  {
    float bias[8];
    const uint4 b0 = lds128(bias_addr), b1 = lds128(bias_addr + 16);
    bias[0] = __uint_as_float(b0.x), bias[1] = __uint_as_float(b0.y), bias[2] = __uint_as_float(b0.z);
    bias[3] = __uint_as_float(b0.w), bias[4] = __uint_as_float(b1.x), bias[5] = __uint_as_float(b1.y);
    bias[6] = __uint_as_float(b1.z), bias[7] = __uint_as_float(b1.w);
    tc_st<8>(t_slot0, bias);
    tc_st<8>(t_slot0 + 32, bias);
    tc_wait_st();
  }
  tc_fence_before();
  __syncwarp();
  if (lane0) mbar_arrive(bar_free);
  uint32_t v0[4], v1[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const f32x2_t x0 = f2_pack(__saturatef(a[2 * i]), __saturatef(a[2 * i + 1]));
    P[PH ^ 1][i] = f2_add(U[i], x0);
    f2_add_to(P[PH][i], P[PH ^ 1][i]);
    U[i] = f2_pack(sat_here(b[2 * i]), sat_here(b[2 * i + 1]));
    Q[PH ^ 1][i] = f2_add(x0, U[i]);
    f2_add_to(Q[PH][i], Q[PH ^ 1][i]);
    if constexpr (SCALE) {
      f32x2_t a2;
      asm volatile("ld.shared.b64 %0, [%1];" : "=l"(a2) : "r"(scale_addr + 8 * i));
      asm("mul.rn.f32x2 %0, %0, %1;" : "+l"(P[PH][i]) : "l"(a2));
      asm("mul.rn.f32x2 %0, %0, %1;" : "+l"(Q[PH][i]) : "l"(a2));
    }
    float lo, hi;
    f2_unpack(P[PH][i], lo, hi);
    v0[i] = HH::pack(lo, hi);
    f2_unpack(Q[PH][i], lo, hi);
    v1[i] = HH::pack(lo, hi);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint32_t t0 = HH::add(v0[i], __shfl_down_sync(0xffffffffu, v0[i], 1));
    const uint32_t t1 = HH::add(v1[i], __shfl_down_sync(0xffffffffu, v1[i], 1));
    hp[0][i] = HH::add(t0, __shfl_down_sync(0xffffffffu, t0, 2));
    hp[1][i] = HH::add(t1, __shfl_down_sync(0xffffffffu, t1, 2));
  }
}

__device__ __forceinline__ void sts128_if(uint32_t addr, const uint32_t (&v)[4], bool pred) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "@p st.shared.v4.b32 [%0], {%1,%2,%3,%4};\n\t"
      "}" ::"r"(addr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(static_cast<uint32_t>(pred))
      : "memory");
}
__device__ __forceinline__ void stg128_if(void* ptr, const uint32_t (&v)[4], bool pred) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "@p st.global.v4.b32 [%0], {%1,%2,%3,%4};\n\t"
      "}" ::"l"(ptr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(static_cast<uint32_t>(pred))
      : "memory");
}

template <bool BF16>
__global__ void __launch_bounds__(kB2Threads, 1) block2_fused_kernel(const B2Params p, const __grid_constant__ B2Maps maps) {
  using Cfg = B2Cfg;
  using HH = H2<BF16>;
  constexpr int R = kB2R, RP = kB2RP, LOGR = kB2LogR;
  constexpr int NS1 = kB2NS1, NS2 = kB2NS2;
  constexpr uint32_t kStageTx = 2 * 4 * 4 * 32 * 16;  // two rows x four planes x four 32-pixel windows

  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* s_w1 = smem;
  uint8_t* s_w2 = s_w1 + Cfg::kWBytes;
  uint8_t* s_st1 = s_w2 + Cfg::kWBytes;
  uint8_t* s_st2 = s_st1 + NS1 * Cfg::kStageBytes;
  uint8_t* s_res = s_st2 + NS2 * Cfg::kStageBytes;
  float* s_bias = reinterpret_cast<float*>(s_res + kB2ResRows * kB2ResRowBytes);  // [2][32]
  uint8_t* s_coef = reinterpret_cast<uint8_t*>(s_bias + 64);  // per channel group: {A x8 fp32 | C x8 fp32 | B x8 16-bit | pad}
  // per residual group (= one epilogue-2 step, two output rows): {ring byte offsets of the upper | lower source row,
  // vertical interpolation weight as a packed 16-bit pair} per row, written by the producer; weight 0xffffffff = no such row
  uint4* s_rdesc = reinterpret_cast<uint4*>(s_coef + 96 * 4);
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_rdesc + kB2ResGroups);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_bar + kB2Bars);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar0 = smem_u32(s_bar);
  const uint32_t bar_l1_full = bar0, bar_l1_empty = bar_l1_full + 8u * NS1;
  const uint32_t bar_l2_full = bar_l1_empty + 8u * NS1, bar_l2_empty = bar_l2_full + 8u * NS2;
  const uint32_t bar_w = bar_l2_empty + 8u * NS2;
  const uint32_t bar_acc1_full = bar_w + 8u, bar_acc1_free = bar_acc1_full + 8u * RP;
  const uint32_t bar_acc2_full = bar_acc1_free + 8u * RP, bar_acc2_free = bar_acc2_full + 8u * RP;
  const uint32_t bar_res_full = bar_acc2_free + 8u * RP, bar_res_done = bar_res_full + 8u * kB2ResGroups;

  if (threadIdx.x == 0) {
    for (int s = 0; s < NS1; ++s) {
      mbar_init(bar_l1_full + 8u * s, 1);
      mbar_init(bar_l1_empty + 8u * s, 1);
    }
    for (int s = 0; s < NS2; ++s) {
      mbar_init(bar_l2_full + 8u * s, 16);  // one arrival per epilogue warp
      mbar_init(bar_l2_empty + 8u * s, 1);
    }
    mbar_init(bar_w, 1);
    for (int s = 0; s < RP; ++s) {
      mbar_init(bar_acc1_full + 8u * s, 1);
      mbar_init(bar_acc1_free + 8u * s, 16);
      mbar_init(bar_acc2_full + 8u * s, 1);
      mbar_init(bar_acc2_free + 8u * s, 16);
    }
    for (int s = 0; s < kB2ResGroups; ++s) {
      mbar_init(bar_res_full + 8u * s, 1);
      mbar_init(bar_res_done + 8u * s, 16);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = threadIdx.x; i < 32; i += kB2Threads) {
    s_bias[i] = p.bias1[i];
    s_bias[32 + i] = p.bias2[i];
  }
  for (int i = threadIdx.x; i < 32; i += kB2Threads) {  // join coefficients (host layout [3][32] fp32: A | B | C)
    uint8_t* g = s_coef + (i >> 3) * 96;
    reinterpret_cast<float*>(g)[i & 7] = p.abc[i];
    reinterpret_cast<float*>(g + 32)[i & 7] = p.abc[64 + i];
    reinterpret_cast<uint16_t*>(g + 64)[i & 7] = static_cast<uint16_t>(HH::pack(p.abc[32 + i], 0.f));  // exact (engine.cu)
  }
  // the 128-byte pad behind the last plane of every stage is read (lanes 126/127, dx taps) but never written
  for (int i = threadIdx.x; i < (NS1 + NS2) * 32; i += kB2Threads)
    reinterpret_cast<uint32_t*>(s_st1 + (i / 32) * Cfg::kStageBytes + 2 * Cfg::kRowBytes)[i % 32] = 0u;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;
  pdl_trigger();
  int row_lo, row_hi;  // this CTA's share of the work line
  rn_gang_rows(blockIdx.x / p.n_strips, gridDim.x / p.n_strips, p.total_rows, p.out_side, p.min_piece, &row_lo, &row_hi);

  if (warp < 4) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kB2RegsCtl));
  if (warp == 0) {
    // ====================== TMA producer: R2 row pairs for layer 1 ======================
    if (lane == 0) {
      mbar_arrive_expect_tx(bar_w, 2 * Cfg::kWBytes);
      for (int off = 0; off < Cfg::kWBytes; off += 9216) {
        tma_bulk_g2s(smem_u32(s_w1 + off), p.w1 + off, 9216, bar_w);
        tma_bulk_g2s(smem_u32(s_w2 + off), p.w2 + off, 9216, bar_w);
      }
      pdl_wait();  // the weights are constants; R2 is the previous kernel's output
      uint32_t st = 0, ph = 1;  // waiting parity 1 on a fresh "empty" barrier passes immediately
      uint32_t G1 = 0;          // global layer-1 conv-row counter (same sequence as the MMA issuer's)
      const uint32_t stage0 = smem_u32(s_st1);
      for (int cur = row_lo; cur < row_hi;) {
        const B2Item it = b2_decode(p, cur, row_hi);
        cur += it.npo;
        const CUtensorMap* tmap = &maps.m[it.strip];
        const int row0 = it.n * p.in_side + it.po0;  // tensor-map row of R2 row 0 of the item
        for (int r = 0; r < it.nin1; r += 2) {
          if (r < it.nconv2) {  // the accumulators this input pair starts must be drained and re-initialised
            const uint32_t gy = G1 + r;
            mbar_wait_sleep(bar_acc1_free + 8u * ((gy >> 1) & (RP - 1)), (gy >> LOGR) & 1);
          }
          mbar_wait_sleep(bar_l1_empty + 8u * st, ph);
          const uint32_t full = bar_l1_full + 8u * st;
          mbar_arrive_expect_tx(full, kStageTx);
          tma_tensor4_g2s(stage0 + st * Cfg::kStageBytes, tmap, 0, 0, 0, row0 + r, full);
          if (++st == NS1) {
            st = 0;
            ph ^= 1;
          }
        }
        G1 += it.nconv2;
      }
    }
  } else if (warp == 2) {
    // ================ TMA producer: residual source rows for the join of layer 2 ================
    // One group = the source rows the two output rows of one epilogue-2 step need that are not in the ring yet.
    // Ring space: three consecutive groups of one item span at most 8 source rows, so group g may be loaded once
    // group g-3 has been released by all epilogue warps; across items, once the previous item is finished.
    if (lane == 0) {
      pdl_wait();
      uint32_t GG = 0;  // global residual-group counter (same sequence as the epilogue's)
      const uint32_t res0 = smem_u32(s_res);
      for (int cur = row_lo; cur < row_hi;) {
        const B2Item it = b2_decode(p, cur, row_hi);
        cur += it.npo;
        // residual window of the strip: columns [jb, jb + 112) of the source rows, one TMA box per row
        const int jb2 = 2 * static_cast<int>(static_cast<float>(p.x0[it.strip]) * p.res_scale);  // in 8-byte elements
        const int row_base = it.n * p.in_side;
        constexpr int kNone = -0x40000000;
        int loaded_hi = kNone;  // highest source row of the item in the ring
        for (int m = 0; m < it.n2e; ++m) {
          const uint32_t g = GG + m;
          if (m >= 3) {
            mbar_wait_sleep(bar_res_done + 8u * ((g - 3) & (kB2ResGroups - 1)), ((g - 3) >> 2) & 1);
          } else if (GG > 0) {
            mbar_wait_sleep(bar_res_done + 8u * ((GG - 1) & (kB2ResGroups - 1)), ((GG - 1) >> 2) & 1);
          }
          const uint32_t full = bar_res_full + 8u * (g & (kB2ResGroups - 1));
          int hi = kNone;
          uint32_t d[4];
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            const int row = 2 * m - 3 + k;
            d[2 * k] = 16u | (16u << 16);  // a row outside the item: harmless ring offsets that match no real row,
            d[2 * k + 1] = 0xffffffffu;    // ... marked by an impossible weight (the stores are predicated off)
            if (row >= 0 && row < it.npo) {  // reference network.py:199 (TF-1.13 legacy bilinear: src = dst * scale)
              const float fy = static_cast<float>(it.po0 + row) * p.res_scale;
              const int y0 = static_cast<int>(fy);
              const int y1 = min(y0 + 1, p.in_side - 1);
              d[2 * k] = static_cast<uint32_t>((y0 & (kB2ResRows - 1)) * kB2ResRowBytes) |
                         (static_cast<uint32_t>((y1 & (kB2ResRows - 1)) * kB2ResRowBytes) << 16);
              d[2 * k + 1] = HH::splat(fy - static_cast<float>(y0));
              if (loaded_hi == kNone) loaded_hi = y0 - 1;  // first row of the item
              hi = y1;
            }
          }
          s_rdesc[g & (kB2ResGroups - 1)] = make_uint4(d[0], d[1], d[2], d[3]);
          if (hi > loaded_hi) {
            mbar_arrive_expect_tx(full, static_cast<uint32_t>(hi - loaded_hi) * kB2ResRowBytes);
            for (int r = loaded_hi + 1; r <= hi; ++r)
              tma_tensor3_g2s(res0 + (r & (kB2ResRows - 1)) * kB2ResRowBytes, &maps.res, jb2, 0, row_base + r, full);
            loaded_hi = hi;
          } else {
            mbar_arrive(full);
          }
        }
        GG += it.n2e;
      }
      // consume the last "group done" phases as well: nothing of this CTA is then left un-waited at exit
      for (uint32_t g = GG > 4 ? GG - 4 : 0; g < GG; ++g)
        mbar_wait_sleep(bar_res_done + 8u * (g & (kB2ResGroups - 1)), (g >> 2) & 1);
    }
  } else if (warp == 1) {
    // ========================= MMA issuer, layer 1 (R2 -> conv2d_2) =========================
    if (elect_one()) {
      mbar_wait(bar_w, 0);
      const uint32_t a_lo0 = (smem_u32(s_st1) >> 4) | (Cfg::kALbo16 << 16);
      const uint32_t b_lo0 = (smem_u32(s_w1) >> 4) | (Cfg::kBLbo16 << 16);
      const uint32_t idesc0 = make_idesc(0, BF16 ? 1 : 0);
      uint32_t st = 0, ph = 0, G = 0;
      for (int cur = row_lo; cur < row_hi;) {
        const B2Item it = b2_decode(p, cur, row_hi);
        cur += it.npo;
        for (int r0 = 0; r0 < it.nin1; r0 += 2) {
          mbar_wait_sleep(bar_l1_full + 8u * st, ph);
          tc_fence_after();
          b2_mma_pair(a_lo0 + st * (Cfg::kStageBytes >> 4), b_lo0, tmem_base, idesc0, G, r0, it.nconv2);
          tc_commit(bar_l1_empty + 8u * st);
          if (r0 >= 2) tc_commit(bar_acc1_full + 8u * (((G + r0 - 2) >> 1) & (RP - 1)));
          if (++st == NS1) {
            st = 0;
            ph ^= 1;
          }
        }
        G += it.nconv2;
      }
    }
    __syncwarp();
  } else if (warp == 3) {
    // ========================= MMA issuer, layer 2 (P2 -> conv2d_3) =========================
    // Its own thread: a P2 stage is multiplied as soon as epilogue 1 has completed it, whatever layer 1 is doing.
    if (elect_one()) {
      mbar_wait(bar_w, 0);
      const uint32_t a_lo0 = (smem_u32(s_st2) >> 4) | (Cfg::kALbo16 << 16);
      const uint32_t b_lo0 = (smem_u32(s_w2) >> 4) | (Cfg::kBLbo16 << 16);
      const uint32_t idesc0 = make_idesc(0, BF16 ? 1 : 0);
      uint32_t st = 0, ph = 0, G = 0;
      for (int cur = row_lo; cur < row_hi;) {
        const B2Item it = b2_decode(p, cur, row_hi);
        cur += it.npo;
        for (int r0 = 0; r0 < it.nin2; r0 += 2) {
          if (r0 < it.nconv3) {  // the accumulators this input pair starts must be drained and re-initialised
            const uint32_t gy = G + r0;
            mbar_wait_sleep(bar_acc2_free + 8u * ((gy >> 1) & (RP - 1)), (gy >> LOGR) & 1);
          }
          mbar_wait_sleep(bar_l2_full + 8u * st, ph);
          tc_fence_after();
          b2_mma_pair(a_lo0 + st * (Cfg::kStageBytes >> 4), b_lo0, tmem_base + R * 32, idesc0, G, r0, it.nconv3);
          tc_commit(bar_l2_empty + 8u * st);
          if (r0 >= 2) tc_commit(bar_acc2_full + 8u * (((G + r0 - 2) >> 1) & (RP - 1)));
          if (++st == NS2) {
            st = 0;
            ph ^= 1;
          }
        }
        G += it.nconv3;
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ============================= epilogue =============================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kB2RegsEpi));
    // (the warp index through a shuffle: ptxas then knows that everything derived from it is warp-uniform and keeps
    // the TMEM / shared-memory base addresses of this warp in uniform registers)
    const int warp_u = __shfl_sync(0xffffffffu, warp, 0);
    const int grp = (warp_u - 4) >> 2;  // channels [8 grp, 8 grp + 8) of both layers
    const int quad = warp_u & 3;        // TMEM lane quadrant this warp may access = window of the 128-pixel tile
    const bool lane0 = lane == 0;
    // Layer-1 lane (window quad, column lane < 27) holds P2 column 27 quad + lane of the strip.  The layer-2 tile is
    // four windows at P2 columns {0, 27, 54, 76}: every P2 column lands in one or two of its lanes.
    const bool l1_lane_ok = lane < 27;
    bool l1_dup = false;
    uint32_t sts_a, sts_b = 0;
    {
      const int ia = quad == 3 ? 101 + lane : 32 * quad + lane;
      int ib = -1;
      if (lane < 5) {
        if (quad == 1) ib = 27 + lane;
        if (quad == 2) ib = 59 + lane;
        if (quad == 3) ib = 91 + lane;
      } else if (quad == 2 && lane >= 22) {
        ib = 74 + lane;
      }
      sts_a = smem_u32(s_st2) + grp * Cfg::kPlaneBytesT + ia * 16;
      l1_dup = l1_lane_ok && ib >= 0;
      if (l1_dup) sts_b = smem_u32(s_st2) + grp * Cfg::kPlaneBytesT + ib * 16;
    }
    // per-lane addresses: ptxas would rather recompute them from %tid in every iteration than keep them in registers:
    // launder them through an opaque move
    asm volatile("mov.b32 %0, %0;" : "+r"(sts_a));
    asm volatile("mov.b32 %0, %0;" : "+r"(sts_b));
    const uint32_t t1_base = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + grp * 8;
    const uint32_t bias_a = smem_u32(s_bias) + grp * 32;  // this thread's 8 biases of layer 1; layer 2: + 128
    const uint32_t coef_a = smem_u32(s_coef) + grp * 96;  // {A x8 fp32 | C x8 fp32 | B x8 16-bit | pad}
    const uint32_t res_ring = smem_u32(s_res) + grp * kB2ResPlaneBytes;
    // layer-2 lane -> output column of the strip
    const int rel2 = (quad == 3 ? 76 : 27 * quad) + lane;
    const bool l2_lane_ok = lane < 27 && (quad < 3 || lane >= 5);
    const uint32_t out_plane_bytes = static_cast<uint32_t>(p.out_side) * 16;
    const uint32_t out_row_bytes = 4 * out_plane_bytes;
    const size_t out_img_bytes = static_cast<size_t>(out_row_bytes) * p.out_side;
    const uint32_t rdesc_a = smem_u32(s_rdesc);

    // every accumulator slot starts out holding the bias: the MMAs then always accumulate
    {
      float b1[8], b2[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        b1[c] = s_bias[grp * 8 + c];
        b2[c] = s_bias[32 + grp * 8 + c];
      }
      for (int s = 0; s < R; ++s) {
        tc_st<8>(t1_base + s * 32, b1);
        tc_st<8>(t1_base + R * 32 + s * 32, b2);
      }
      tc_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane0)
        for (int s = 0; s < RP; ++s) {
          mbar_arrive(bar_acc1_free + 8u * s);
          mbar_arrive(bar_acc2_free + 8u * s);
        }
    }
    pdl_wait();  // before the first global store

    // ring positions, advanced incrementally (all rings have four entries): accumulator pair + parity of the two
    // layers, P2 stage + parity of the stage the even row of an epilogue-1 step starts, residual group + parity
    uint32_t a1 = 0, a1_par = 0, a2 = 0, a2_par = 0, sg = 0, sg_par = 1;
    for (int cur = row_lo; cur < row_hi;) {
      const B2Item it = b2_decode(p, cur, row_hi);
      cur += it.npo;
      const int x0 = p.x0[it.strip];
      const int col = x0 + rel2;
      const bool col_ok = l2_lane_ok && col >= p.own_lo[it.strip] && col < p.own_hi[it.strip];
      // residual taps of this thread's output column (reference network.py:199, TF-1.13 legacy bilinear)
      const int jb = static_cast<int>(static_cast<float>(x0) * p.res_scale);
      const float fx = static_cast<float>(min(col, p.out_side - 1)) * p.res_scale;
      const int jx0 = static_cast<int>(fx);
      const uint32_t jtx2 = HH::splat(fx - static_cast<float>(jx0));
      const uint32_t jl = res_ring + static_cast<uint32_t>(min(jx0 - jb, kB2ResPx - 2)) * 16;  // left tap, ring row 0
      // (right tap = jl + 16: with a resize scale > 1 the last output column still has a right neighbour)
      uint32_t J[4] = {0u, 0u, 0u, 0u}, Jn[4];  // J: horizontally interpolated source row at ring offset `joff`
      uint32_t joff = 0xffffffffu;
      // output row 2m - 3 of epilogue-2 step m (advanced by two rows per step; starts three rows above the item)
      uint8_t* optr = p.out + it.n * out_img_bytes + static_cast<size_t>(it.po0) * out_row_bytes + grp * out_plane_bytes +
                      static_cast<size_t>(min(col, p.out_side - 1)) * 16 - 3 * static_cast<ptrdiff_t>(out_row_bytes);

      f32x2_t U1[4], P1[2][4], Q1[2][4], U2[4], P2[2][4], Q2[2][4];  // vertical windows of the two layers (b2_drain_pool)
#pragma unroll
      for (int c = 0; c < 4; ++c) U1[c] = P1[0][c] = Q1[0][c] = U2[c] = P2[0][c] = Q2[0][c] = 0ull;

      // ---------------- layer 1: conv rows (2t, 2t+1) -> P2 rows (2t-3, 2t-2) into the layer-2 stages ----------
      auto layer1_step = [&](auto ph, int t) {
        mbar_wait_sleep(bar_acc1_full + 8u * a1, a1_par);
        tc_fence_after();
        uint32_t hp[2][4];
        b2_drain_pool<HH, decltype(ph)::value, false>(t1_base + a1 * 64, bias_a, bar_acc1_free + 8u * a1, lane0, U1, P1, Q1,
                                                      hp);
        a1 = (a1 + 1) & 3;
        a1_par ^= a1 == 0;
        // P2 row 2t-3 completes the stage of pair t-2 (= the stage before `sg`), row 2t-2 starts pair t-1 in `sg`
        if (t >= 2) {
          const uint32_t dst = ((sg + 3) & 3) * Cfg::kStageBytes + Cfg::kRowBytes;
          sts128_if(sts_a + dst, hp[0], l1_lane_ok);
          sts128_if(sts_b + dst, hp[0], l1_dup);
          // generic-proxy writes -> visible to the tensor core's async-proxy reads, then signal the MMA issuer
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (lane0) mbar_arrive(bar_l2_full + 8u * ((sg + 3) & 3));
        }
        if (t >= 1 && t <= it.n2s) {  // row 2t-2 < nin2
          mbar_wait_sleep(bar_l2_empty + 8u * sg, sg_par);  // the stage must have been consumed by the layer-2 MMAs
          const uint32_t dst = sg * Cfg::kStageBytes;
          sts128_if(sts_a + dst, hp[1], l1_lane_ok);
          sts128_if(sts_b + dst, hp[1], l1_dup);
          sg = (sg + 1) & 3;
          sg_par ^= sg == 0;
        }
      };
      // ---------------- layer 2: conv rows (2m, 2m+1) -> output rows (2m-3, 2m-2) + residual join ---------------
      auto layer2_step = [&](auto ph) {
        mbar_wait_sleep(bar_acc2_full + 8u * a2, a2_par);
        tc_fence_after();
        uint32_t hp[2][4];
        b2_drain_pool<HH, decltype(ph)::value, true>(t1_base + R * 32 + a2 * 64, bias_a + 128, bar_acc2_free + 8u * a2, lane0,
                                                     U2, P2, Q2, hp, coef_a);
        // (the residual-group ring advances in lockstep with the layer-2 accumulator ring: same index, same parity)
        const uint32_t rg = a2, rg_par = a2_par;
        a2 = (a2 + 1) & 3;
        a2_par ^= a2 == 0;
        mbar_wait_sleep(bar_res_full + 8u * rg, rg_par);
        // reference network.py:199-203 in folded form: bilinear taps in packed 16-bit arithmetic
        // (top = tl + (tr - tl) * tx, ... : the TF formula).  Which ring rows and which vertical weight: the
        // producer's descriptor of this group (one broadcast load); a row outside the item has weight 0xffffffff.
        const uint4 rd = lds128(rdesc_a + 16u * rg);
        uint32_t rs[2][4];
        const bool row_ok0 = rd.y != 0xffffffffu, row_ok1 = rd.w != 0xffffffffu;
        auto hlerp = [&](uint32_t (&dst)[4], uint32_t off) {
          const uint4 l = lds128(jl + off), r = lds128(jl + off + 16);
          dst[0] = HH::fma(HH::sub(r.x, l.x), jtx2, l.x);
          dst[1] = HH::fma(HH::sub(r.y, l.y), jtx2, l.y);
          dst[2] = HH::fma(HH::sub(r.z, l.z), jtx2, l.z);
          dst[3] = HH::fma(HH::sub(r.w, l.w), jtx2, l.w);
        };
        {  // first row: upper source row in J (normally left there by the previous output row), lower one -> Jn
          const uint32_t o0 = rd.x & 0xffffu, o1 = rd.x >> 16;
          if (o0 != joff) hlerp(J, o0);
          hlerp(Jn, o1);
#pragma unroll
          for (int i = 0; i < 4; ++i) rs[0][i] = HH::fma(HH::sub(Jn[i], J[i]), rd.y, J[i]);
          joff = o1;
        }
        {  // second row: the roles of J and Jn swap, so the lower source row ends up in J again
          const uint32_t o0 = rd.z & 0xffffu, o1 = rd.z >> 16;
          if (o0 != joff) hlerp(Jn, o0);
          hlerp(J, o1);
#pragma unroll
          for (int i = 0; i < 4; ++i) rs[1][i] = HH::fma(HH::sub(J[i], Jn[i]), rd.w, Jn[i]);
          joff = o1;
        }
        // the ring rows of this group are in registers: hand the group back to the producer
        __syncwarp();
        if (lane0) mbar_arrive(bar_res_done + 8u * rg);
        // out = A * pool + (B * resized + C): A was applied to the fp32 window sums, B is a 16-bit value by construction
        // (engine.cu stores the channel with a gain that makes it one) and multiplies the 16-bit resized residual
        // in a mixed-precision fma with fp32 accumulation; the pooled 16-bit value joins through a mixed-precision add
        {
          const uint4 C0 = lds128(coef_a + 32), C1 = lds128(coef_a + 48), Bh = lds128(coef_a + 64);
          const uint32_t Bv[4] = {Bh.x, Bh.y, Bh.z, Bh.w};
          const float Cv[8] = {__uint_as_float(C0.x), __uint_as_float(C0.y), __uint_as_float(C0.z), __uint_as_float(C0.w),
                               __uint_as_float(C1.x), __uint_as_float(C1.y), __uint_as_float(C1.z), __uint_as_float(C1.w)};
#pragma unroll
          for (int k = 0; k < 2; ++k)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float lo = HH::fhadd_lo(hp[k][i], HH::fhfma_lo(rs[k][i], Bv[i], Cv[2 * i]));
              const float hi = HH::fhadd_hi(hp[k][i], HH::fhfma_hi(rs[k][i], Bv[i], Cv[2 * i + 1]));
              hp[k][i] = HH::pack(lo, hi);
            }
        }
        stg128_if(optr, hp[0], col_ok && row_ok0);
        stg128_if(optr + out_row_bytes, hp[1], col_ok && row_ok1);
        optr += 2 * out_row_bytes;
      };

      // Layer 1 runs kB2LagEpi = 5 row pairs ahead of layer 2 (n1e = n2e + 3 >= 5).  The window phases alternate
      // per call; every item starts in phase 0 and the step sequence is arranged so that they are compile-time:
      //   5 x L1 | pairs of {L1<1> L2<0> L1<0> L2<1>} | [L1<1>] L2<0> L2<1> [L2<0>]   (bracketed if n2e is odd)
      constexpr std::integral_constant<int, 0> ph0{};
      constexpr std::integral_constant<int, 1> ph1{};
      static_assert(kB2LagEpi == 5, "the phase schedule below is written for a lag of five row pairs");
      layer1_step(ph0, 0);
      layer1_step(ph1, 1);
      layer1_step(ph0, 2);
      layer1_step(ph1, 3);
      layer1_step(ph0, 4);
      int t = kB2LagEpi;
      for (; t + 2 <= it.n1e; t += 2) {  // steady state
        layer1_step(ph1, t);
        layer2_step(ph0);
        layer1_step(ph0, t + 1);
        layer2_step(ph1);
      }
      const bool odd = t < it.n1e;
      if (odd) layer1_step(ph1, t);
      layer2_step(ph0);
      layer2_step(ph1);
      if (odd) layer2_step(ph0);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

}  // namespace

bool Block2FusedSupported(const TcConvLayer& l1, const TcConvLayer& l2) {
  if (l1.cin != 32 || l1.cout != 32 || l2.cin != 32 || l2.cout != 32) return false;
  if (l1.pool_k != 4 || l1.pool_s != 1 || l2.pool_k != 4 || l2.pool_s != 1) return false;
  if (l1.cout_parts != 1 || l2.cout_parts != 1 || l1.amode != 0 || l2.amode != 0) return false;
  if (l1.out_side != l2.in_side || l2.join_src_side != l1.in_side || !l2.join_abc) return false;
  if (l2.out_side < kB2StripOut) return false;  // narrower maps: the layer-by-layer kernels
  if ((l2.out_side + kB2StripOut - 1) / kB2StripOut > kB2MaxStrips) return false;
  // residual window of a strip / of three consecutive row groups must fit the shared-memory ring
  const float scale = static_cast<float>(l1.in_side) / static_cast<float>(l2.out_side);
  if (l1.in_side <= l2.out_side || scale * (kB2StripOut - 1) + 3.f > static_cast<float>(kB2ResPx)) return false;
  if (scale * 5.f + 2.f > static_cast<float>(kB2ResRows)) return false;
  return true;
}

cudaError_t Block2Fused(const TcConvLayer& l1, const TcConvLayer& l2, const void* in, void* out, int N, HalfKind kind,
                        cudaStream_t st) {
  if (!Block2FusedSupported(l1, l2)) return cudaErrorInvalidValue;
  B2Params p{};
  p.in = static_cast<const uint8_t*>(in);
  p.out = static_cast<uint8_t*>(out);
  p.w1 = static_cast<const uint8_t*>(l1.w_packed);
  p.w2 = static_cast<const uint8_t*>(l2.w_packed);
  p.bias1 = l1.bias;
  p.bias2 = l2.bias;
  p.abc = l2.join_abc;
  p.N = N;
  p.in_side = l1.in_side;
  p.out_side = l2.out_side;
  p.res_scale = static_cast<float>(l1.in_side) / static_cast<float>(l2.out_side);
  p.n_strips = (p.out_side + kB2StripOut - 1) / kB2StripOut;
  for (int k = 0; k < p.n_strips; ++k) {
    p.x0[k] = std::min(k * kB2StripOut, p.out_side - kB2StripOut);
    p.own_lo[k] = k * kB2StripOut;
    p.own_hi[k] = std::min((k + 1) * kB2StripOut, p.out_side);
  }
  // every gang of n_strips CTAs gets the same share of the (image x output row) line (tc_common.cuh)
  p.total_rows = N * p.out_side;
  int gangs = 1;
  rn_plan_rows(p.total_rows, p.out_side, p.n_strips, std::max(p.n_strips, SmCount()), &gangs, &p.min_piece);
  const int gx = gangs * p.n_strips;

  const PFN_encodeTiled encode = GetEncodeTiled();
  if (!encode) return cudaErrorNotSupported;
  B2Maps maps;
  std::memset(&maps, 0, sizeof(maps));
  for (int k = 0; k < p.n_strips; ++k) {
    // dims (fastest first): 256 elements = one 32-pixel window | window (stride 27 pixels, overlapping) |
    // channel-chunk plane | image row.  Box = four windows x four planes x a pair of rows.
    const cuuint64_t gdim[4] = {256, 4, 4, static_cast<cuuint64_t>(N) * p.in_side};
    const cuuint64_t gstr[3] = {27 * 16, static_cast<cuuint64_t>(p.in_side) * 16,
                                static_cast<cuuint64_t>(4) * p.in_side * 16};
    const cuuint32_t box[4] = {256, 4, 4, 2};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult cr = encode(&maps.m[k], CU_TENSOR_MAP_DATA_TYPE_UINT16, 4,
                         const_cast<uint8_t*>(p.in) + static_cast<size_t>(p.x0[k]) * 16, gdim, gstr, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) return cudaErrorInvalidValue;
  }
  {
    // residual rows: 8-byte elements so that one box row holds 112 pixels (a 16-bit box is limited to 256 elements)
    const cuuint64_t gdim[3] = {static_cast<cuuint64_t>(2) * p.in_side, 4, static_cast<cuuint64_t>(N) * p.in_side};
    const cuuint64_t gstr[2] = {static_cast<cuuint64_t>(p.in_side) * 16, static_cast<cuuint64_t>(4) * p.in_side * 16};
    const cuuint32_t box[3] = {2 * kB2ResPx, 4, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    CUresult cr = encode(&maps.res, CU_TENSOR_MAP_DATA_TYPE_UINT64, 3, const_cast<uint8_t*>(p.in), gdim, gstr, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) return cudaErrorInvalidValue;
  }
  auto kern = kind == HalfKind::kBF16 ? block2_fused_kernel<true> : block2_fused_kernel<false>;
  cudaError_t ea = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kB2SmemBytes);
  if (ea != cudaSuccess) return ea;
  cudaError_t el = LaunchPdl(kern, dim3(gx), dim3(kB2Threads), kB2SmemBytes, st, N, p, maps);
  if (el != cudaSuccess) return el;
  return cudaGetLastError();
}

}  // namespace rn
