// Device half of the JPEG front end (cv2.imread of infer.py:81 for baseline JPEG files): quantised DCT coefficients ->
// BGR uint8 image, with the integer arithmetic of the decoder cv2 links (libjpeg-turbo defaults: "slow integer"
// inverse DCT with 13-bit constants, triangle-filter ("fancy") chroma upsampling, 16-bit fixed-point YCbCr -> RGB), so
// the result is bit-identical to cv2.imread's and the logits downstream are too.  Both kernels are HBM-bound byte /
// integer work: coefficients are read once (2 B per sample), sample planes written and read once, 3 B per pixel out.
#include "kernels.h"

namespace rn {
namespace {

// ---- inverse DCT (Loeffler-Ligtenberg-Moschytz, 2 passes, CONST_BITS = 13, PASS1_BITS = 2) ----
constexpr int kF0_298 = 2446, kF0_390 = 3196, kF0_541 = 4433, kF0_765 = 6270, kF0_899 = 7373, kF1_175 = 9633,
              kF1_501 = 12299, kF1_847 = 15137, kF1_961 = 16069, kF2_053 = 16819, kF2_562 = 20995, kF3_072 = 25172;

template <int SHIFT>
__device__ __forceinline__ int descale(int x) {
  return (x + (1 << (SHIFT - 1))) >> SHIFT;
}

// one 8-point pass; in[] are the eight inputs, out[] the eight outputs before the descale
__device__ __forceinline__ void idct8(const int in[8], int out[8]) {
  int z2 = in[2], z3 = in[6];
  int z1 = (z2 + z3) * kF0_541;
  int tmp2 = z1 - z3 * kF1_847;
  int tmp3 = z1 + z2 * kF0_765;
  z2 = in[0];
  z3 = in[4];
  int tmp0 = (z2 + z3) << 13;
  int tmp1 = (z2 - z3) << 13;
  const int tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
  tmp0 = in[7];
  tmp1 = in[5];
  tmp2 = in[3];
  tmp3 = in[1];
  z1 = tmp0 + tmp3;
  z2 = tmp1 + tmp2;
  z3 = tmp0 + tmp2;
  int z4 = tmp1 + tmp3;
  const int z5 = (z3 + z4) * kF1_175;
  tmp0 *= kF0_298;
  tmp1 *= kF2_053;
  tmp2 *= kF3_072;
  tmp3 *= kF1_501;
  z1 *= -kF0_899;
  z2 *= -kF2_562;
  z3 = z3 * -kF1_961 + z5;
  z4 = z4 * -kF0_390 + z5;
  tmp0 += z1 + z3;
  tmp1 += z2 + z4;
  tmp2 += z2 + z3;
  tmp3 += z1 + z4;
  out[0] = tmp10 + tmp3;
  out[7] = tmp10 - tmp3;
  out[1] = tmp11 + tmp2;
  out[6] = tmp11 - tmp2;
  out[2] = tmp12 + tmp1;
  out[5] = tmp12 - tmp1;
  out[3] = tmp13 + tmp0;
  out[4] = tmp13 - tmp0;
}

// the decoder's post-IDCT range-limit table, indexed with 10 bits of (value): centre +128, clamp, wrap beyond +-512
__device__ __forceinline__ unsigned range_limit(int v) {
  const int x = v & 1023;
  return x < 128 ? 128 + x : (x < 512 ? 255 : (x < 896 ? 0 : x - 896));
}

constexpr int kIdctThreads = 256;                 // 8 threads per 8x8 block
constexpr int kBlocksPerIter = kIdctThreads / 8;  // 32 blocks per CUDA block and iteration

__global__ void __launch_bounds__(kIdctThreads)
jpeg_idct_kernel(const int16_t* __restrict__ coefs, const JpegPlaneDesc* __restrict__ planes,
                 const uint16_t* __restrict__ quant, uint8_t* __restrict__ samples) {
  __shared__ int ws[kBlocksPerIter][8][9];
  __shared__ int s_q[64];
  const JpegPlaneDesc d = planes[blockIdx.y];
  if (threadIdx.x < 64) s_q[threadIdx.x] = quant[d.quant_index * 64 + threadIdx.x];
  __syncthreads();
  const int jb = threadIdx.x >> 3, c = threadIdx.x & 7;
  const int nblocks = d.wblocks * d.hblocks;
  const int16_t* cbase = coefs + d.coef_offset;
  uint8_t* pbase = samples + d.plane_offset;
  const int pitch = d.wblocks * 8;
  for (int b0 = blockIdx.x * kBlocksPerIter; b0 < nblocks; b0 += gridDim.x * kBlocksPerIter) {
    const int b = b0 + jb;
    if (b < nblocks) {
      // pass 1: column c of the dequantised block
      int in[8], out[8];
      const int16_t* cp = cbase + static_cast<size_t>(b) * 64 + c;
#pragma unroll
      for (int r = 0; r < 8; ++r) in[r] = static_cast<int>(cp[r * 8]) * s_q[r * 8 + c];
      idct8(in, out);
#pragma unroll
      for (int r = 0; r < 8; ++r) ws[jb][r][c] = descale<13 - 2>(out[r]);
    }
    __syncwarp();
    if (b < nblocks) {
      // pass 2: row c of the intermediate block -> eight samples
      int in[8], out[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) in[k] = ws[jb][c][k];
      idct8(in, out);
      unsigned lo = 0, hi = 0;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        lo |= range_limit(descale<13 + 2 + 3>(out[k])) << (8 * k);
        hi |= range_limit(descale<13 + 2 + 3>(out[4 + k])) << (8 * k);
      }
      const int by = b / d.wblocks, bx = b - by * d.wblocks;
      *reinterpret_cast<uint2*>(pbase + static_cast<size_t>(by * 8 + c) * pitch + bx * 8) = make_uint2(lo, hi);
    }
    __syncwarp();
  }
}

// ---- chroma upsampling + colour conversion + EXIF orientation -> packed BGR ----
__device__ __forceinline__ int clamp255(int v) { return min(max(v, 0), 255); }

// Four neighbouring chroma samples a b c d (columns i-1 .. i+2; a / d clamped at the plane border) -> the four pixels
// 2i .. 2i+3.  Triangle filter of the decoder: 3/4 nearer + 1/4 further sample, rounding 1 for even and 2 for odd
// outputs; the first and last column of the plane pass through unfiltered.
__device__ __forceinline__ void up_h2v1(int a, int b, int c, int d, int i, int dw, int out[4]) {
  out[0] = i == 0 ? b : (3 * b + a + 1) >> 2;
  out[1] = i == dw - 1 ? b : (3 * b + c + 2) >> 2;
  out[2] = (3 * c + b + 1) >> 2;
  out[3] = i + 1 == dw - 1 ? c : (3 * c + d + 2) >> 2;
}
// Both directions (9/16, 3/16, 3/16, 1/16): a .. d are the column sums 3 * nearer row + further row, rounding 8 / 7.
__device__ __forceinline__ void up_h2v2(int a, int b, int c, int d, int i, int dw, int out[4]) {
  out[0] = i == 0 ? (4 * b + 8) >> 4 : (3 * b + a + 8) >> 4;
  out[1] = i == dw - 1 ? (4 * b + 7) >> 4 : (3 * b + c + 7) >> 4;
  out[2] = (3 * c + b + 8) >> 4;
  out[3] = i + 1 == dw - 1 ? (4 * c + 7) >> 4 : (3 * c + d + 7) >> 4;
}

constexpr int kColorThreads = 256;
constexpr int kColorTile = kColorThreads * 4;  // pixels of one row handled by a block per iteration

// One thread = four neighbouring pixels of a row: one 32-bit load of luma, four (eight for 4:2:0) byte loads per chroma
// plane, three 32-bit stores (the oriented image has a row pitch that is a multiple of four pixels).
__global__ void __launch_bounds__(kColorThreads)
jpeg_color_kernel(const uint8_t* __restrict__ samples, const JpegImageDesc* __restrict__ images, uint8_t* __restrict__ raw) {
  const JpegImageDesc d = images[blockIdx.y];
  const int W = d.width, H = d.height;
  const uint8_t* py = samples + d.plane[0];
  const uint8_t* pcb = samples + d.plane[1];
  const uint8_t* pcr = samples + d.plane[2];
  uint8_t* out = raw + d.out_offset;
  const int tiles_per_row = (W + kColorTile - 1) / kColorTile;
  const int total = H * tiles_per_row;
  for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
    const int y = tile / tiles_per_row;
    const int x = (tile - y * tiles_per_row) * kColorTile + threadIdx.x * 4;
    if (x >= W) continue;
    const unsigned yw = *reinterpret_cast<const unsigned*>(py + static_cast<size_t>(y) * d.pitch[0] + x);
    int cb[4] = {128, 128, 128, 128}, cr[4] = {128, 128, 128, 128};
    if (d.ncomp == 3) {
      if (d.hs == 1) {
        const unsigned bw = *reinterpret_cast<const unsigned*>(pcb + static_cast<size_t>(y) * d.pitch[1] + x);
        const unsigned rw = *reinterpret_cast<const unsigned*>(pcr + static_cast<size_t>(y) * d.pitch[2] + x);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          cb[k] = (bw >> (8 * k)) & 255;
          cr[k] = (rw >> (8 * k)) & 255;
        }
      } else {
        const int i = x >> 1, dw = d.cdw;
        const int ia = max(i - 1, 0), ic = min(i + 1, dw - 1), id = min(i + 2, dw - 1);
        if (d.vs == 1) {
          const uint8_t* rb = pcb + static_cast<size_t>(y) * d.pitch[1];
          const uint8_t* rr = pcr + static_cast<size_t>(y) * d.pitch[2];
          up_h2v1(rb[ia], rb[i], rb[ic], rb[id], i, dw, cb);
          up_h2v1(rr[ia], rr[i], rr[ic], rr[id], i, dw, cr);
        } else {
          // nearer chroma row y / 2, further row above (even y) or below (odd y), replicated at the border
          const int r0 = y >> 1;
          const int r1 = (y & 1) ? min(r0 + 1, d.cdh - 1) : max(r0 - 1, 0);
          const uint8_t* b0 = pcb + static_cast<size_t>(r0) * d.pitch[1];
          const uint8_t* b1 = pcb + static_cast<size_t>(r1) * d.pitch[1];
          const uint8_t* q0 = pcr + static_cast<size_t>(r0) * d.pitch[2];
          const uint8_t* q1 = pcr + static_cast<size_t>(r1) * d.pitch[2];
          up_h2v2(3 * b0[ia] + b1[ia], 3 * b0[i] + b1[i], 3 * b0[ic] + b1[ic], 3 * b0[id] + b1[id], i, dw, cb);
          up_h2v2(3 * q0[ia] + q1[ia], 3 * q0[i] + q1[i], 3 * q0[ic] + q1[ic], 3 * q0[id] + q1[id], i, dw, cr);
        }
      }
    }
    unsigned char px[12];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int Y = (yw >> (8 * k)) & 255;
      const int u = cb[k] - 128, v = cr[k] - 128;
      // FIX(1.40200) = 91881, FIX(1.77200) = 116130, FIX(0.71414) = 46802, FIX(0.34414) = 22554 at 16 fractional bits
      px[3 * k + 0] = static_cast<unsigned char>(clamp255(Y + ((116130 * u + 32768) >> 16)));
      px[3 * k + 1] = static_cast<unsigned char>(clamp255(Y + ((-22554 * u + 32768 - 46802 * v) >> 16)));
      px[3 * k + 2] = static_cast<unsigned char>(clamp255(Y + ((91881 * v + 32768) >> 16)));
    }
    if (d.orientation == 1 && x + 4 <= W) {
      unsigned* o = reinterpret_cast<unsigned*>(out + (static_cast<size_t>(y) * d.out_pitch + x) * 3);
      o[0] = px[0] | (px[1] << 8) | (px[2] << 16) | (px[3] << 24);
      o[1] = px[4] | (px[5] << 8) | (px[6] << 16) | (px[7] << 24);
      o[2] = px[8] | (px[9] << 8) | (px[10] << 16) | (px[11] << 24);
      continue;
    }
    // row tail, or an EXIF orientation as cv2.imread applies it (transpose and / or flips of the decoded image)
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int xx = x + k;
      if (xx >= W) break;
      int ox = xx, oy = y;
      switch (d.orientation) {
        case 2: ox = W - 1 - xx; break;
        case 3: ox = W - 1 - xx; oy = H - 1 - y; break;
        case 4: oy = H - 1 - y; break;
        case 5: ox = y; oy = xx; break;
        case 6: ox = H - 1 - y; oy = xx; break;
        case 7: ox = H - 1 - y; oy = W - 1 - xx; break;
        case 8: ox = y; oy = W - 1 - xx; break;
        default: break;
      }
      uint8_t* o = out + (static_cast<size_t>(oy) * d.out_pitch + ox) * 3;
      o[0] = px[3 * k + 0];
      o[1] = px[3 * k + 1];
      o[2] = px[3 * k + 2];
    }
  }
}

}  // namespace

cudaError_t JpegIdct(const int16_t* coefs, const JpegPlaneDesc* planes, int n_planes, const uint16_t* quant,
                     uint8_t* samples, cudaStream_t st) {
  if (n_planes <= 0) return cudaSuccess;
  // a photograph has tens of thousands of blocks per plane: 4 CUDA blocks per SM in x, every plane in y
  dim3 grid(148 * 4 / (n_planes < 4 ? n_planes : 4) + 1, n_planes);
  jpeg_idct_kernel<<<grid, kIdctThreads, 0, st>>>(coefs, planes, quant, samples);
  return cudaGetLastError();
}

cudaError_t JpegColor(const uint8_t* samples, const JpegImageDesc* images, int n_images, uint8_t* raw, cudaStream_t st) {
  if (n_images <= 0) return cudaSuccess;
  dim3 grid(148 * 8 / (n_images < 8 ? n_images : 8) + 1, n_images);
  jpeg_color_kernel<<<grid, kColorThreads, 0, st>>>(samples, images, raw);
  return cudaGetLastError();
}

}  // namespace rn
