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
// Warp roles: warp 0 = TMA producer (one thread), warp 1 = MMA issuer (one elected thread, both layers, statically
// interleaved), warps 2..17 = epilogue: warp (quadrant q, channel group g) owns TMEM lanes 32q..32q+31 and
// channels 8g..8g+7 of BOTH layers and alternates between them (layer 2 runs kLagEpi row pairs behind layer 1).
#include <cstring>

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
constexpr int kB2LagMma = 5;          // MMA issuer: layer-2 row pair j is issued together with layer-1 pair j + 5
constexpr int kB2LagEpi = 5;          // epilogue warps: layer-2 pair m is drained after layer-1 pair m + 5
constexpr int kB2LagRes = 7;          // producer: residual group m is requested together with R2 pair m + 7
constexpr int kB2Threads = 64 + 16 * 32;
constexpr int kB2Bars = 2 * kB2NS1 + 2 * kB2NS2 + 1 + 4 * kB2RP + 2 * kB2ResGroups;
constexpr int kB2SmemBytes = 2 * B2Cfg::kWBytes + (kB2NS1 + kB2NS2) * B2Cfg::kStageBytes + kB2ResRows * kB2ResRowBytes +
                             (2 * 32 + 3 * 32) * 4 + kB2Bars * 8 + 16;
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
  int n_strips, rows_per_item, n_rowblocks, n_items;
  float res_scale;     // in_side / out_side as float32 (TF computes the resize scale in float32)
  int x0[kB2MaxStrips];      // first column a strip computes (input, P2 and output columns share the origin)
  int own_lo[kB2MaxStrips];  // output columns [own_lo, own_hi) are stored by the strip
  int own_hi[kB2MaxStrips];
};

struct B2Maps {
  CUtensorMap m[kB2MaxStrips];  // per strip: R2 windows {256 el | 4 windows, stride 27 px | 4 planes | N*in_side rows}
};

struct B2Item {
  int n, strip, po0, npo;
  int nconv3, nin2, nconv2, nin1;  // conv rows / input rows of the two layers (all even)
  int n1s, n1e, n2s, n2e;          // R2 stage loads, epilogue-1 steps, P2 pairs, epilogue-2 steps
};

__device__ __forceinline__ B2Item b2_decode(const B2Params& p, int item) {
  B2Item it;
  const int rb = item % p.n_rowblocks;
  const int t = item / p.n_rowblocks;
  it.strip = t % p.n_strips;
  it.n = t / p.n_strips;
  it.po0 = rb * p.rows_per_item;
  it.npo = min(p.rows_per_item, p.out_side - it.po0);
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

// Drain one accumulator row pair (8 channels of this thread's pixel), hand the slots back pre-loaded with the bias,
// clip (saturate = ReLU6/6, the factor lives in the weights), advance the vertical 4-row window (fp32 registers) and
// apply the horizontal 4-column window on packed 16-bit pairs with warp shuffles.
// hp[0] = pooled row (y - 3), hp[1] = pooled row (y - 2) for conv rows (y, y + 1) of this call.
template <typename HH>
__device__ __forceinline__ void b2_drain_pool(uint32_t t_base, uint32_t slot0, uint32_t slot1, const float* s_bias8,
                                              uint32_t bar_free, int lane, float (&r1)[8], float (&q1)[8], float (&q2)[8],
                                              uint32_t (&hp)[2][4]) {
  float a[8], b[8];
  tc_ld<8>(t_base + slot0 * 32, a);
  tc_ld<8>(t_base + slot1 * 32, b);
  tc_wait_ld();
  {
    float bias[8];
    const float4 b0 = *reinterpret_cast<const float4*>(s_bias8);
    const float4 b1 = *reinterpret_cast<const float4*>(s_bias8 + 4);
    bias[0] = b0.x, bias[1] = b0.y, bias[2] = b0.z, bias[3] = b0.w;
    bias[4] = b1.x, bias[5] = b1.y, bias[6] = b1.z, bias[7] = b1.w;
    tc_st<8>(t_base + slot0 * 32, bias);
    tc_st<8>(t_base + slot1 * 32, bias);
    tc_wait_st();
  }
  tc_fence_before();
  __syncwarp();
  if (lane == 0) mbar_arrive(bar_free);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float o0[2], o1[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int c = 2 * i + e;
      const float x0 = __saturatef(a[c]), x1 = __saturatef(b[c]);
      const float qa = x0 + r1[c], qb = x1 + x0;
      o0[e] = qa + q2[c];
      o1[e] = qb + q1[c];
      q2[c] = qa;
      q1[c] = qb;
      r1[c] = x1;
    }
    const uint32_t v0 = HH::pack(o0[0], o0[1]), v1 = HH::pack(o1[0], o1[1]);
    const uint32_t t0 = HH::add(v0, __shfl_down_sync(0xffffffffu, v0, 1));
    hp[0][i] = HH::add(t0, __shfl_down_sync(0xffffffffu, t0, 2));
    const uint32_t t1 = HH::add(v1, __shfl_down_sync(0xffffffffu, v1, 1));
    hp[1][i] = HH::add(t1, __shfl_down_sync(0xffffffffu, t1, 2));
  }
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
  float* s_abc = s_bias + 64;                                                     // [3][32]
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_abc + 96);
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
  for (int i = threadIdx.x; i < 96; i += kB2Threads) s_abc[i] = p.abc[i];
  // the 128-byte pad behind the last plane of every stage is read (lanes 126/127, dx taps) but never written
  for (int i = threadIdx.x; i < (NS1 + NS2) * 32; i += kB2Threads)
    reinterpret_cast<uint32_t*>(s_st1 + (i / 32) * Cfg::kStageBytes + 2 * Cfg::kRowBytes)[i % 32] = 0u;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;
  pdl_trigger();

  const size_t in_plane_bytes = static_cast<size_t>(p.in_side) * 16;
  const size_t in_row_bytes = 4 * in_plane_bytes;
  const size_t in_img_bytes = in_row_bytes * p.in_side;

  if (warp == 0) {
    // =========================== TMA producer ===========================
    if (lane == 0) {
      mbar_arrive_expect_tx(bar_w, 2 * Cfg::kWBytes);
      for (int off = 0; off < Cfg::kWBytes; off += 9216) {
        tma_bulk_g2s(smem_u32(s_w1 + off), p.w1 + off, 9216, bar_w);
        tma_bulk_g2s(smem_u32(s_w2 + off), p.w2 + off, 9216, bar_w);
      }
      pdl_wait();  // the weights are constants; R2 is the previous kernel's output
      uint32_t st = 0, ph = 1;  // waiting parity 1 on a fresh "empty" barrier passes immediately
      uint32_t G1 = 0;          // global layer-1 conv-row counter (same sequence as the MMA issuer's)
      uint32_t GG = 0;          // global residual-group counter (same sequence as the epilogue's)
      const uint32_t stage0 = smem_u32(s_st1), res0 = smem_u32(s_res);
      for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
        const B2Item it = b2_decode(p, item);
        const CUtensorMap* tmap = &maps.m[it.strip];
        const int row0 = it.n * p.in_side + it.po0;  // tensor-map row of R2 row 0 of the item
        // residual window of the strip: columns [jb, jb + 112) of the source rows
        const int jb = static_cast<int>(static_cast<float>(p.x0[it.strip]) * p.res_scale);
        const uint8_t* res_src = p.in + it.n * in_img_bytes + static_cast<size_t>(jb) * 16;
        int loaded_hi = -1;
        const int steps = max(it.n1s, it.n2e + kB2LagRes);
        for (int s = 0; s < steps; ++s) {
          if (s < it.n1s) {
            const int r = 2 * s;
            if (r < it.nconv2) {  // the accumulators this input pair starts must be drained and re-initialised
              const uint32_t gy = G1 + r;
              mbar_wait(bar_acc1_free + 8u * ((gy >> 1) & (RP - 1)), (gy >> LOGR) & 1);
            }
            mbar_wait(bar_l1_empty + 8u * st, ph);
            const uint32_t full = bar_l1_full + 8u * st;
            mbar_arrive_expect_tx(full, kStageTx);
            tma_tensor4_g2s(stage0 + st * Cfg::kStageBytes, tmap, 0, 0, 0, row0 + r, full);
            if (++st == NS1) {
              st = 0;
              ph ^= 1;
            }
          }
          const int m = s - kB2LagRes;
          if (m >= 0 && m < it.n2e) {
            const uint32_t g = GG + m;
            // ring space: three consecutive groups of one item span at most 8 source rows; across items wait for
            // the previous item's last group.  With the static offset kB2LagRes both are normally long complete.
            if (m >= 3) {
              mbar_wait(bar_res_done + 8u * ((g - 3) & (kB2ResGroups - 1)), ((g - 3) >> 2) & 1);
            } else if (GG > 0) {
              mbar_wait(bar_res_done + 8u * ((GG - 1) & (kB2ResGroups - 1)), ((GG - 1) >> 2) & 1);
            }
            const int ra = max(2 * m - 3, 0), rb = min(2 * m - 2, it.npo - 1);
            const uint32_t full = bar_res_full + 8u * (g & (kB2ResGroups - 1));
            if (rb >= ra) {
              const int lo = static_cast<int>(static_cast<float>(it.po0 + ra) * p.res_scale);
              const int hi = min(static_cast<int>(static_cast<float>(it.po0 + rb) * p.res_scale) + 1, p.in_side - 1);
              const int first = loaded_hi < 0 ? lo : loaded_hi + 1;
              const int nrows = hi - first + 1;
              if (nrows > 0) {
                mbar_arrive_expect_tx(full, static_cast<uint32_t>(nrows) * kB2ResRowBytes);
                for (int r = first; r <= hi; ++r) {
                  const uint8_t* src = res_src + static_cast<size_t>(r) * in_row_bytes;
                  const uint32_t dst = res0 + (r & (kB2ResRows - 1)) * kB2ResRowBytes;
#pragma unroll
                  for (int c = 0; c < 4; ++c)
                    tma_bulk_g2s(dst + c * kB2ResPlaneBytes, src + c * in_plane_bytes, kB2ResPlaneBytes, full);
                }
                loaded_hi = hi;
              } else {
                mbar_arrive(full);
              }
            } else {
              mbar_arrive(full);
            }
          }
        }
        G1 += it.nconv2;
        GG += it.n2e;
      }
    }
  } else if (warp == 1) {
    // ============================ MMA issuer ============================
    if (elect_one()) {
      mbar_wait(bar_w, 0);
      const uint32_t a1_lo0 = (smem_u32(s_st1) >> 4) | (Cfg::kALbo16 << 16);
      const uint32_t a2_lo0 = (smem_u32(s_st2) >> 4) | (Cfg::kALbo16 << 16);
      const uint32_t b1_lo0 = (smem_u32(s_w1) >> 4) | (Cfg::kBLbo16 << 16);
      const uint32_t b2_lo0 = (smem_u32(s_w2) >> 4) | (Cfg::kBLbo16 << 16);
      const uint32_t idesc0 = make_idesc(0, BF16 ? 1 : 0);
      uint32_t st1 = 0, ph1 = 0, st2 = 0, ph2 = 0, G1 = 0, G2 = 0;
      for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
        const B2Item it = b2_decode(p, item);
        const int steps = max(it.n1s, it.n2s + kB2LagMma);
        for (int s = 0; s < steps; ++s) {
          if (s < it.n1s) {
            const int r0 = 2 * s;
            mbar_wait(bar_l1_full + 8u * st1, ph1);
            tc_fence_after();
            b2_mma_pair(a1_lo0 + st1 * (Cfg::kStageBytes >> 4), b1_lo0, tmem_base, idesc0, G1, r0, it.nconv2);
            tc_commit(bar_l1_empty + 8u * st1);
            if (r0 >= 2) tc_commit(bar_acc1_full + 8u * (((G1 + r0 - 2) >> 1) & (RP - 1)));
            if (++st1 == NS1) {
              st1 = 0;
              ph1 ^= 1;
            }
          }
          const int j = s - kB2LagMma;
          if (j >= 0 && j < it.n2s) {
            const int r0 = 2 * j;
            if (r0 < it.nconv3) {
              const uint32_t gy = G2 + r0;
              mbar_wait(bar_acc2_free + 8u * ((gy >> 1) & (RP - 1)), (gy >> LOGR) & 1);
            }
            mbar_wait(bar_l2_full + 8u * st2, ph2);
            tc_fence_after();
            b2_mma_pair(a2_lo0 + st2 * (Cfg::kStageBytes >> 4), b2_lo0, tmem_base + R * 32, idesc0, G2, r0, it.nconv3);
            tc_commit(bar_l2_empty + 8u * st2);
            if (r0 >= 2) tc_commit(bar_acc2_full + 8u * (((G2 + r0 - 2) >> 1) & (RP - 1)));
            if (++st2 == NS2) {
              st2 = 0;
              ph2 ^= 1;
            }
          }
        }
        G1 += it.nconv2;
        G2 += it.nconv3;
      }
    }
    __syncwarp();
  } else {
    // ============================= epilogue =============================
    const int grp = (warp - 2) >> 2;  // channels [8 grp, 8 grp + 8) of both layers
    const int quad = warp & 3;        // TMEM lane quadrant this warp may access = window of the 128-pixel tile
    const uint32_t t1_base = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + grp * 8;
    const uint32_t t2_base = t1_base + R * 32;
    const float* s_bias1 = s_bias + grp * 8;
    const float* s_bias2 = s_bias + 32 + grp * 8;
    // Layer-1 lane (window quad, column lane < 27) holds P2 column 27 quad + lane of the strip.  The layer-2 tile is
    // four windows at P2 columns {0, 27, 54, 76}: every P2 column lands in one or two of its lanes.
    uint32_t sts_a, sts_b = 0xffffffffu;
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
      sts_a = grp * Cfg::kPlaneBytesT + ia * 16;
      if (ib >= 0) sts_b = grp * Cfg::kPlaneBytesT + ib * 16;
    }
    const bool l1_lane_ok = lane < 27;
    // layer-2 lane -> output column of the strip
    const int rel2 = (quad == 3 ? 76 : 27 * quad) + lane;
    const bool l2_lane_ok = lane < 27 && (quad < 3 || lane >= 5);
    const size_t out_plane_bytes = static_cast<size_t>(p.out_side) * 16;
    const size_t out_row_bytes = 4 * out_plane_bytes;
    const size_t out_img_bytes = out_row_bytes * p.out_side;

    // every accumulator slot starts out holding the bias: the MMAs then always accumulate
    {
      float b1[8], b2[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        b1[c] = s_bias1[c];
        b2[c] = s_bias2[c];
      }
      for (int s = 0; s < R; ++s) {
        tc_st<8>(t1_base + s * 32, b1);
        tc_st<8>(t2_base + s * 32, b2);
      }
      tc_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0)
        for (int s = 0; s < RP; ++s) {
          mbar_arrive(bar_acc1_free + 8u * s);
          mbar_arrive(bar_acc2_free + 8u * s);
        }
    }
    pdl_wait();  // before the first global store

    uint32_t G1 = 0, G2 = 0, J2 = 0, GG = 0;
    const uint32_t st2_0 = smem_u32(s_st2);
    const uint8_t* res_ring = s_res + grp * kB2ResPlaneBytes;
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
      const B2Item it = b2_decode(p, item);
      const int x0 = p.x0[it.strip];
      const int col = x0 + rel2;
      const bool col_ok = l2_lane_ok && col >= p.own_lo[it.strip] && col < p.own_hi[it.strip];
      // residual taps of this thread's output column (reference network.py:199, TF-1.13 legacy bilinear)
      const int jb = static_cast<int>(static_cast<float>(x0) * p.res_scale);
      const float fx = static_cast<float>(min(col, p.out_side - 1)) * p.res_scale;
      const int jx0 = static_cast<int>(fx);
      const uint32_t jtx2 = HH::splat(fx - static_cast<float>(jx0));
      const uint32_t joff = static_cast<uint32_t>(min(jx0 - jb, kB2ResPx - 2)) * 16;
      const uint32_t jdx = jx0 + 1 < p.in_side ? 16u : 0u;
      uint32_t jbot[4] = {0u, 0u, 0u, 0u};
      int jy_prev = -1;
      uint8_t* optr = p.out + it.n * out_img_bytes + static_cast<size_t>(it.po0) * out_row_bytes + grp * out_plane_bytes +
                      static_cast<size_t>(min(col, p.out_side - 1)) * 16;

      float r1a[8], q1a[8], q2a[8], r1b[8], q1b[8], q2b[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) r1a[c] = q1a[c] = q2a[c] = r1b[c] = q1b[c] = q2b[c] = 0.f;

      const int steps = max(it.n1e, it.n2e + kB2LagEpi);
      for (int t = 0; t < steps; ++t) {
        if (t < it.n1e) {
          // ---------------- layer 1: conv rows (2t, 2t+1) -> P2 rows (2t-3, 2t-2) into the layer-2 stages ----
          const uint32_t gy = G1 + 2 * t;
          const uint32_t pair = (gy >> 1) & (RP - 1);
          mbar_wait(bar_acc1_full + 8u * pair, (gy >> LOGR) & 1);
          tc_fence_after();
          uint32_t hp[2][4];
          b2_drain_pool<HH>(t1_base, gy & (R - 1), (gy + 1) & (R - 1), s_bias1, bar_acc1_free + 8u * pair, lane, r1a, q1a,
                            q2a, hp);
          const int row_odd = 2 * t - 3, row_even = 2 * t - 2;
          if (row_even >= 0 && row_even < it.nin2) {  // first row of P2 pair t-1: its stage must have been consumed
            const uint32_t jg = J2 + (t - 1);
            const uint32_t stg = jg & (NS2 - 1);
            mbar_wait(bar_l2_empty + 8u * stg, ((jg / NS2) & 1) ^ 1);
            const uint32_t dst = st2_0 + stg * Cfg::kStageBytes;
            if (l1_lane_ok) {
              asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(dst + sts_a), "r"(hp[1][0]), "r"(hp[1][1]),
                           "r"(hp[1][2]), "r"(hp[1][3])
                           : "memory");
              if (sts_b != 0xffffffffu)
                asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(dst + sts_b), "r"(hp[1][0]), "r"(hp[1][1]),
                             "r"(hp[1][2]), "r"(hp[1][3])
                             : "memory");
            }
          }
          if (row_odd >= 0 && row_odd < it.nin2) {  // second row of P2 pair t-2: completes the stage
            const uint32_t jg = J2 + (t - 2);
            const uint32_t stg = jg & (NS2 - 1);
            const uint32_t dst = st2_0 + stg * Cfg::kStageBytes + Cfg::kRowBytes;
            if (l1_lane_ok) {
              asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(dst + sts_a), "r"(hp[0][0]), "r"(hp[0][1]),
                           "r"(hp[0][2]), "r"(hp[0][3])
                           : "memory");
              if (sts_b != 0xffffffffu)
                asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(dst + sts_b), "r"(hp[0][0]), "r"(hp[0][1]),
                             "r"(hp[0][2]), "r"(hp[0][3])
                             : "memory");
            }
            // generic-proxy writes -> visible to the tensor core's async-proxy reads, then signal the MMA issuer
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_l2_full + 8u * stg);
          }
        }
        const int m = t - kB2LagEpi;
        if (m >= 0 && m < it.n2e) {
          // ---------------- layer 2: conv rows (2m, 2m+1) -> output rows (2m-3, 2m-2) + residual join ---------
          const uint32_t gy = G2 + 2 * m;
          const uint32_t pair = (gy >> 1) & (RP - 1);
          mbar_wait(bar_acc2_full + 8u * pair, (gy >> LOGR) & 1);
          tc_fence_after();
          uint32_t hp[2][4];
          b2_drain_pool<HH>(t2_base, gy & (R - 1), (gy + 1) & (R - 1), s_bias2, bar_acc2_free + 8u * pair, lane, r1b, q1b,
                            q2b, hp);
          const uint32_t g = GG + m;
          mbar_wait(bar_res_full + 8u * (g & (kB2ResGroups - 1)), (g >> 2) & 1);
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            const int row = 2 * m - 3 + k;
            if (row >= 0 && row < it.npo) {
              // reference network.py:199-203 in folded form: bilinear taps in packed 16-bit arithmetic
              // (top = tl + (tr - tl) * tx, ... : the TF formula), the per-channel affine in fp32
              const float fy = static_cast<float>(it.po0 + row) * p.res_scale;
              const int y0 = static_cast<int>(fy);
              const int y1 = min(y0 + 1, p.in_side - 1);
              const uint32_t ty2 = HH::splat(fy - static_cast<float>(y0));
              uint32_t top[4];
              if (y0 == jy_prev) {
#pragma unroll
                for (int i = 0; i < 4; ++i) top[i] = jbot[i];
              } else {
                const uint8_t* a = res_ring + (y0 & (kB2ResRows - 1)) * kB2ResRowBytes + joff;
                const uint4 l = *reinterpret_cast<const uint4*>(a);
                const uint4 r = *reinterpret_cast<const uint4*>(a + jdx);
                top[0] = HH::fma(HH::sub(r.x, l.x), jtx2, l.x);
                top[1] = HH::fma(HH::sub(r.y, l.y), jtx2, l.y);
                top[2] = HH::fma(HH::sub(r.z, l.z), jtx2, l.z);
                top[3] = HH::fma(HH::sub(r.w, l.w), jtx2, l.w);
              }
              {
                const uint8_t* a = res_ring + (y1 & (kB2ResRows - 1)) * kB2ResRowBytes + joff;
                const uint4 l = *reinterpret_cast<const uint4*>(a);
                const uint4 r = *reinterpret_cast<const uint4*>(a + jdx);
                jbot[0] = HH::fma(HH::sub(r.x, l.x), jtx2, l.x);
                jbot[1] = HH::fma(HH::sub(r.y, l.y), jtx2, l.y);
                jbot[2] = HH::fma(HH::sub(r.z, l.z), jtx2, l.z);
                jbot[3] = HH::fma(HH::sub(r.w, l.w), jtx2, l.w);
              }
              jy_prev = y1;
              uint32_t o[4];
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float2 rs = HH::unpack(HH::fma(HH::sub(jbot[i], top[i]), ty2, top[i]));
                const float2 hv = HH::unpack(hp[k][i]);
                const float* co = s_abc + grp * 8 + 2 * i;
                const float2 a2 = *reinterpret_cast<const float2*>(co);
                const float2 b2 = *reinterpret_cast<const float2*>(co + 32);
                const float2 c2 = *reinterpret_cast<const float2*>(co + 64);
                o[i] = HH::pack(fmaf(a2.x, hv.x, fmaf(b2.x, rs.x, c2.x)), fmaf(a2.y, hv.y, fmaf(b2.y, rs.y, c2.y)));
              }
              if (col_ok)
                *reinterpret_cast<uint4*>(optr + static_cast<size_t>(row) * out_row_bytes) = make_uint4(o[0], o[1], o[2], o[3]);
            }
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_res_done + 8u * (g & (kB2ResGroups - 1)));
        }
      }
      G1 += it.nconv2;
      G2 += it.nconv3;
      J2 += it.n2s;
      GG += it.n2e;
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
  if (scale < 1.f || scale * (kB2StripOut - 1) + 3.f > static_cast<float>(kB2ResPx)) return false;
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
  // Row blocks: items are dealt round-robin to the persistent CTAs; pick the split that minimises
  // (items per CTA, rounded up) x (row pairs per item + halo rows of the two layers + pipeline fill).
  {
    const int ctas = std::max(1, SmCount());
    const int max_nrb = std::max(1, p.out_side / 8);
    long best_cost = -1;
    int best_rows = p.out_side;
    for (int nrb = 1; nrb <= max_nrb; ++nrb) {
      const int rows = (p.out_side + nrb - 1) / nrb;
      const int blocks = (p.out_side + rows - 1) / rows;
      const long items = static_cast<long>(N) * p.n_strips * blocks;
      const long rounds = (items + ctas - 1) / ctas;
      const long cost = rounds * (rows + 10 + 2 * kB2LagEpi);
      if (best_cost < 0 || cost < best_cost) {
        best_cost = cost;
        best_rows = rows;
      }
    }
    p.rows_per_item = best_rows;
    p.n_rowblocks = (p.out_side + best_rows - 1) / best_rows;
  }
  p.n_items = N * p.n_strips * p.n_rowblocks;

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
  auto kern = kind == HalfKind::kBF16 ? block2_fused_kernel<true> : block2_fused_kernel<false>;
  cudaError_t ea = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kB2SmemBytes);
  if (ea != cudaSuccess) return ea;
  const int gx = std::max(1, std::min(p.n_items, SmCount()));
  cudaError_t el = LaunchPdl(kern, dim3(gx), dim3(kB2Threads), kB2SmemBytes, st, N, p, maps);
  if (el != cudaSuccess) return el;
  return cudaGetLastError();
}

}  // namespace rn
