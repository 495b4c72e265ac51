// FP32 CUDA-core kernels (RN_PREC_FP32 path and the small-channel front/tail of
// the 16-bit path).  Layout: NHWC fp32 activations, HWIO folded weights.
//
// Folded layer form:  p' = pool(relu6(conv_W(p) + b))
//   conv3x3_relu6_kernel : reference network.py:184-186 (conv2d VALID + ReLU6)
//   avgpool_kernel       : reference network.py:188-191 (avg_pool VALID)
//   join_kernel          : reference network.py:199-203 (+resize_bilinear, BN) as A*p+B*resize(p0)+C
//   dense_tail_kernel    : reference network.py:210-223, :231-237, :44-45
#include "kernels.h"

namespace rn {

int SmCount() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

namespace {

__device__ __forceinline__ float relu6f(float v) { return fminf(fmaxf(v, 0.f), 6.f); }

// ---------------------------------------------------------------------------
// Direct 3x3 VALID convolution + bias + ReLU6.
// Block = 16x16 output pixels, each thread owns one pixel and COG output channels.
// Input channels are consumed in chunks of CC through shared memory:
//   tile [CC][18][18+1]  (channel-major so that a warp's x-consecutive reads hit
//                         consecutive banks), weights [9][CC][COG] (warp-broadcast).
// ---------------------------------------------------------------------------
constexpr int kTile = 16;
constexpr int kHalo = kTile + 2;

template <typename TIn>
__device__ __forceinline__ float load_as_float(const TIn* p) {
  return static_cast<float>(*p);
}

template <int CC, int COG, typename TIn>
__global__ void __launch_bounds__(kTile* kTile)
    conv3x3_relu6_kernel(const TIn* __restrict__ in, const float* __restrict__ w, const float* __restrict__ bias,
                         float* __restrict__ out, int H, int W, int Cin, int Cout, int px_stride) {
  __shared__ float s_in[CC][kHalo][kHalo + 1];
  __shared__ __align__(16) float s_w[9][CC][COG];

  const int OH = H - 2, OW = W - 2;
  const int groups = Cout / COG;
  const int n = blockIdx.z / groups;
  const int co0 = (blockIdx.z % groups) * COG;
  const int tx = threadIdx.x % kTile, ty = threadIdx.x / kTile;
  const int oy0 = blockIdx.y * kTile, ox0 = blockIdx.x * kTile;
  const TIn* in_n = in + static_cast<size_t>(n) * H * W * px_stride;  // px_stride > Cin: padded pixels (BGRA)

  float acc[COG];
#pragma unroll
  for (int o = 0; o < COG; ++o) acc[o] = 0.f;

  for (int c0 = 0; c0 < Cin; c0 += CC) {
    // stage the input halo tile: consecutive threads walk (x, c) so global reads stay contiguous
    for (int idx = threadIdx.x; idx < kHalo * kHalo * CC; idx += kTile * kTile) {
      int c = idx % CC;
      int px = (idx / CC) % kHalo;
      int py = idx / (CC * kHalo);
      int iy = oy0 + py, ix = ox0 + px;
      float v = 0.f;
      if (iy < H && ix < W && c0 + c < Cin) v = load_as_float(in_n + (static_cast<size_t>(iy) * W + ix) * px_stride + c0 + c);
      s_in[c][py][px] = v;
    }
    for (int idx = threadIdx.x; idx < 9 * CC * COG; idx += kTile * kTile) {
      int o = idx % COG;
      int c = (idx / COG) % CC;
      int tap = idx / (COG * CC);
      float v = 0.f;
      if (c0 + c < Cin) v = w[(static_cast<size_t>(tap) * Cin + c0 + c) * Cout + co0 + o];
      s_w[tap][c][o] = v;
    }
    __syncthreads();
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      const int dy = tap / 3, dx = tap % 3;
#pragma unroll
      for (int c = 0; c < CC; ++c) {
        const float a = s_in[c][ty + dy][tx + dx];
#pragma unroll
        for (int o = 0; o < COG; ++o) acc[o] = fmaf(a, s_w[tap][c][o], acc[o]);
      }
    }
    __syncthreads();
  }
  const int oy = oy0 + ty, ox = ox0 + tx;
  if (oy < OH && ox < OW) {
    float* o_ptr = out + ((static_cast<size_t>(n) * OH + oy) * OW + ox) * Cout + co0;
#pragma unroll
    for (int o = 0; o < COG; ++o) o_ptr[o] = relu6f(acc[o] + bias[co0 + o]);
  }
}

// k x k / stride s VALID average pooling, one thread per output element (c fastest).
__global__ void avgpool_kernel(const float* __restrict__ in, float* __restrict__ out, int N, int H, int W, int C,
                               int k, int s, int OH, int OW) {
  const size_t total = static_cast<size_t>(N) * OH * OW * C;
  const float inv = 1.f / static_cast<float>(k * k);
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    int c = static_cast<int>(i % C);
    size_t r = i / C;
    int ox = static_cast<int>(r % OW);
    r /= OW;
    int oy = static_cast<int>(r % OH);
    int n = static_cast<int>(r / OH);
    const float* p = in + ((static_cast<size_t>(n) * H + oy * s) * W + ox * s) * C + c;
    float acc = 0.f;
    for (int dy = 0; dy < k; ++dy)
      for (int dx = 0; dx < k; ++dx) acc += p[(static_cast<size_t>(dy) * W + dx) * C];
    out[i] = acc * inv;
  }
}

// out = A*p + B*resize_bilinear_legacy(src) + C   (TF-1.13 ResizeBilinear, align_corners=False)
__global__ void join_kernel(const float* __restrict__ p, const float* __restrict__ src, float* __restrict__ out,
                            const float* __restrict__ A, const float* __restrict__ B, const float* __restrict__ Cc,
                            int N, int S, int SS, int C) {
  const size_t total = static_cast<size_t>(N) * S * S * C;
  const float scale = static_cast<float>(SS) / static_cast<float>(S);
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    int c = static_cast<int>(i % C);
    size_t r = i / C;
    int x = static_cast<int>(r % S);
    r /= S;
    int y = static_cast<int>(r % S);
    int n = static_cast<int>(r / S);
    float fy = static_cast<float>(y) * scale, fx = static_cast<float>(x) * scale;
    int y0 = static_cast<int>(floorf(fy)), x0 = static_cast<int>(floorf(fx));
    int y1 = min(y0 + 1, SS - 1), x1 = min(x0 + 1, SS - 1);
    float ty = fy - static_cast<float>(y0), tx = fx - static_cast<float>(x0);
    const float* s_n = src + static_cast<size_t>(n) * SS * SS * C + c;
    float tl = s_n[(static_cast<size_t>(y0) * SS + x0) * C], tr = s_n[(static_cast<size_t>(y0) * SS + x1) * C];
    float bl = s_n[(static_cast<size_t>(y1) * SS + x0) * C], br = s_n[(static_cast<size_t>(y1) * SS + x1) * C];
    float top = tl + (tr - tl) * tx;
    float bot = bl + (br - bl) * tx;
    float res = top + (bot - top) * ty;
    out[i] = fmaf(A[c], p[i], fmaf(B[c], res, Cc[c]));
  }
}

// Dense head: 4 x (matmul + bias + ReLU6), softmax, argmax.  One warp per image.
// Layer widths after the first are <= 32 so a lane owns one output unit.
__global__ void dense_tail_kernel(const float* __restrict__ flat, int N, int flat_len, DenseParams dp,
                                  long long* __restrict__ top1, float* __restrict__ probs,
                                  float* __restrict__ logits, float* __restrict__ pre_relu6) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (warp >= N) return;
  const float* x = flat + static_cast<size_t>(warp) * flat_len;
  // layer 0: flat_len -> n0 (<=32): lane o accumulates its column in index order (deterministic)
  float v = 0.f;
  {
    const int on = dp.out[0];
    if (lane < on) {
      float acc = 0.f;
      for (int r = 0; r < flat_len; ++r) acc = fmaf(x[r], dp.w[0][static_cast<size_t>(r) * on + lane], acc);
      v = relu6f(acc + dp.b[0][lane]);
    }
  }
  float pre = 0.f;
#pragma unroll
  for (int l = 1; l < 4; ++l) {
    const int in = dp.out[l - 1], on = dp.out[l];
    float acc = 0.f;
    for (int r = 0; r < in; ++r) {
      float xr = __shfl_sync(0xffffffffu, v, r);
      if (lane < on) acc = fmaf(xr, dp.w[l][r * on + lane], acc);
    }
    if (lane < on) acc += dp.b[l][lane];
    pre = acc;
    v = relu6f(acc);
  }
  const int C = dp.out[3];
  // softmax over the ReLU6-clipped logits (reference network.py:43-44), argmax of the softmax (:45)
  float m = lane < C ? v : -INFINITY;
  for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  float e = lane < C ? expf(v - m) : 0.f;
  float s = e;
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  float prob = e / s;
  // first index of the maximum probability
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
    if (probs) probs[static_cast<size_t>(warp) * C + lane] = prob;
    if (logits) logits[static_cast<size_t>(warp) * C + lane] = v;
    if (pre_relu6) pre_relu6[static_cast<size_t>(warp) * C + lane] = pre;
  }
  if (lane == 0 && top1) top1[warp] = bi;
}

// Centre crop + OpenCV-compatible fixed-point bilinear resize of one uint8 HxWx3 image (reference
// network.py:137-146 + cv2.resize at :152).  taps[0..3] = x0, x1, a0, a1 (per output column), taps[4..7] = y0, y1,
// b0, b1 (per output row), all int32; arithmetic = OpenCV's 11-bit coefficient path:
//   h = S[x0]*a0 + S[x1]*a1 ;  out = (((b0*(h0>>4))>>16) + ((b1*(h1>>4))>>16) + 2) >> 2
// area2x: exact 2x down-scaling in both directions is the 2x2 box filter (a+b+c+d+2)>>2, as OpenCV does.
__global__ void crop_resize_u8_kernel(const uint8_t* __restrict__ src, int W, int cy, int cx, uint8_t* __restrict__ dst,
                                      int S, const int* __restrict__ taps, int area2x) {
  const int dx = blockIdx.x * blockDim.x + threadIdx.x, dy = blockIdx.y;
  if (dx >= S) return;
  uint8_t* o = dst + (static_cast<size_t>(dy) * S + dx) * 3;
  if (area2x) {
    const uint8_t* p0 = src + (static_cast<size_t>(cy + 2 * dy) * W + cx + 2 * dx) * 3;
    const uint8_t* p1 = p0 + static_cast<size_t>(W) * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) o[c] = static_cast<uint8_t>((p0[c] + p0[3 + c] + p1[c] + p1[3 + c] + 2) >> 2);
    return;
  }
  const int x0 = taps[dx], x1 = taps[S + dx], a0 = taps[2 * S + dx], a1 = taps[3 * S + dx];
  const int y0 = taps[4 * S + dy], y1 = taps[5 * S + dy], b0 = taps[6 * S + dy], b1 = taps[7 * S + dy];
  const uint8_t* r0 = src + (static_cast<size_t>(cy + y0) * W + cx) * 3;
  const uint8_t* r1 = src + (static_cast<size_t>(cy + y1) * W + cx) * 3;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const int h0 = r0[x0 * 3 + c] * a0 + r0[x1 * 3 + c] * a1;
    const int h1 = r1[x0 * 3 + c] * a0 + r1[x1 * 3 + c] * a1;
    const int v = (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2;
    o[c] = static_cast<uint8_t>(min(max(v, 0), 255));
  }
}

// Batched form for a list of photos of different sizes (classify_im_dir, reference infer.py:79-82): one launch crops and
// resizes every image of the micro-batch from a device arena straight into the network's input tensor.  The tap tables
// of crop_resize_u8_kernel are evaluated in place with OpenCV's own expressions (resize.cpp, double / float steps as
// in engine.cu: ResizeTaps), so the result is the same bit pattern: fx = float((d + 0.5) * scale - 0.5), floor,
// 11-bit weights by round-half-even; columns clamp the fraction at the borders, rows only clamp the indices.
__global__ void crop_resize_batch_kernel(const uint8_t* __restrict__ arena, const CropDesc* __restrict__ descs,
                                         uint8_t* __restrict__ dst, int S) {
  const int dx = blockIdx.x * blockDim.x + threadIdx.x, dy = blockIdx.y;
  if (dx >= S) return;
  const CropDesc d = descs[blockIdx.z];
  const uint8_t* src = arena + d.offset;
  uint8_t* o = dst + ((static_cast<size_t>(blockIdx.z) * S + dy) * S + dx) * 3;
  const int W = d.W, cy = d.cy, cx = d.cx, side = d.side;
  if (side == S) {  // already the right size: plain crop (network.py:151 skips the resize)
    const uint8_t* p0 = src + (static_cast<size_t>(cy + dy) * W + cx + dx) * 3;
    o[0] = p0[0], o[1] = p0[1], o[2] = p0[2];
    return;
  }
  if (side == 2 * S) {  // exact 2x shrink: 2x2 box
    const uint8_t* p0 = src + (static_cast<size_t>(cy + 2 * dy) * W + cx + 2 * dx) * 3;
    const uint8_t* p1 = p0 + static_cast<size_t>(W) * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) o[c] = static_cast<uint8_t>((p0[c] + p0[3 + c] + p1[c] + p1[3 + c] + 2) >> 2);
    return;
  }
  const double scale = static_cast<double>(side) / static_cast<double>(S);
  float fx = static_cast<float>((dx + 0.5) * scale - 0.5);
  int sx = static_cast<int>(floorf(fx));
  fx -= static_cast<float>(sx);
  if (sx < 0) {
    fx = 0.f;
    sx = 0;
  }
  if (sx >= side - 1) {
    fx = 0.f;
    sx = side - 1;
  }
  const int x0 = sx, x1 = min(sx + 1, side - 1);
  const int a1 = static_cast<int>(rintf(fx * 2048.0f)), a0 = static_cast<int>(rintf((1.0f - fx) * 2048.0f));
  float fy = static_cast<float>((dy + 0.5) * scale - 0.5);
  const int sy = static_cast<int>(floorf(fy));
  fy -= static_cast<float>(sy);
  const int y0 = min(max(sy, 0), side - 1), y1 = min(max(sy + 1, 0), side - 1);
  const int b1 = static_cast<int>(rintf(fy * 2048.0f)), b0 = static_cast<int>(rintf((1.0f - fy) * 2048.0f));
  const uint8_t* r0 = src + (static_cast<size_t>(cy + y0) * W + cx) * 3;
  const uint8_t* r1 = src + (static_cast<size_t>(cy + y1) * W + cx) * 3;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const int h0 = r0[x0 * 3 + c] * a0 + r0[x1 * 3 + c] * a1;
    const int h1 = r1[x0 * 3 + c] * a0 + r1[x1 * 3 + c] * a1;
    const int v = (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2;
    o[c] = static_cast<uint8_t>(min(max(v, 0), 255));
  }
}

// Camera front end of the mobile module folded into the device call (reference ImageUtils.java:131-151
// convertYUV420ToARGB8888 + :88-129 YUV2RGB, ClassifierActivity.java:89-106 frameToCropTransform + drawBitmap):
// one thread per network input pixel maps its centre back through the frame-to-crop transform (rotation by a multiple
// of 90 degrees about the frame centre, uniform scale max(S/inW, S/inH), see getTransformationMatrix :168-225),
// takes the nearest frame pixel (an unfiltered Canvas.drawBitmap), converts it with the reference's integer YUV -> RGB
// arithmetic and writes the R,G,B bytes of the uint8 RGB feed.  The coordinate arithmetic is fp64 with a fixed
// operation order (oracle/yuv_front.py restates it with the same order).
__global__ void yuv420_crop_kernel(const uint8_t* __restrict__ yp, const uint8_t* __restrict__ up,
                                   const uint8_t* __restrict__ vp, YuvFrame f, uint8_t* __restrict__ dst, int S) {
  const int dx = blockIdx.x * blockDim.x + threadIdx.x, dy = blockIdx.y;
  if (dx >= S) return;
  const bool transpose = (f.rotation + 90) % 180 == 0;
  const int in_w = transpose ? f.height : f.width, in_h = transpose ? f.width : f.height;
  double sx = 1.0, sy = 1.0;
  if (in_w != S || in_h != S) {
    const double fx = static_cast<double>(S) / in_w, fy = static_cast<double>(S) / in_h;
    sx = sy = fx > fy ? fx : fy;  // MAINTAIN_ASPECT (ClassifierActivity.java:40)
  }
  // inverse of  T(S/2) * Scale * R(rotation) * T(-frame/2)   (without the translations when rotation == 0)
  double px = dx + 0.5, py = dy + 0.5;
  if (f.rotation != 0) {
    px -= S / 2.0;
    py -= S / 2.0;
  }
  px /= sx;
  py /= sy;
  double qx = px, qy = py;
  if (f.rotation == 90) {  // forward: (x, y) -> (-y, x)
    qx = py;
    qy = -px;
  } else if (f.rotation == 180) {
    qx = -px;
    qy = -py;
  } else if (f.rotation == 270) {  // forward: (x, y) -> (y, -x)
    qx = -py;
    qy = px;
  }
  if (f.rotation != 0) {
    qx += f.width / 2.0;
    qy += f.height / 2.0;
  }
  const int i = min(max(static_cast<int>(floor(qx)), 0), f.width - 1);
  const int j = min(max(static_cast<int>(floor(qy)), 0), f.height - 1);
  int y = yp[f.y_row_stride * j + i];
  const int uv = f.uv_row_stride * (j >> 1) + (i >> 1) * f.uv_pixel_stride;
  int u = up[uv], v = vp[uv];
  y = (y - 16) < 0 ? 0 : (y - 16);
  u -= 128;
  v -= 128;
  const int y1192 = 1192 * y;
  int r = y1192 + 1634 * v, g = y1192 - 833 * v - 400 * u, b = y1192 + 2066 * u;
  r = min(max(r, 0), 262143);
  g = min(max(g, 0), 262143);
  b = min(max(b, 0), 262143);
  uint8_t* o = dst + (static_cast<size_t>(dy) * S + dx) * 3;
  o[0] = static_cast<uint8_t>(r >> 10);  // 0xff000000 | ((r << 6) & 0xff0000) | ((g >> 2) & 0xff00) | ((b >> 10) & 0xff)
  o[1] = static_cast<uint8_t>(g >> 10);
  o[2] = static_cast<uint8_t>(b >> 10);
}

template <int CC, int COG, typename TIn>
void launch_conv(const TIn* in, const float* w, const float* b, float* out, int N, int H, int W, int Cin, int Cout,
                 int px_stride, cudaStream_t st) {
  dim3 grid((W - 2 + kTile - 1) / kTile, (H - 2 + kTile - 1) / kTile, N * (Cout / COG));
  conv3x3_relu6_kernel<CC, COG, TIn><<<grid, kTile * kTile, 0, st>>>(in, w, b, out, H, W, Cin, Cout, px_stride);
}

}  // namespace

template <typename TIn>
cudaError_t Conv3x3Relu6F32(const TIn* in, const float* w, const float* b, float* out, int N, int H, int W, int Cin,
                            int Cout, cudaStream_t st, int px_stride) {
  if (px_stride <= 0) px_stride = Cin;
  if (Cin == 3 && Cout % 8 == 0)
    launch_conv<3, 8, TIn>(in, w, b, out, N, H, W, Cin, Cout, px_stride, st);
  else if (Cin % 8 == 0 && Cout % 16 == 0)
    launch_conv<8, 16, TIn>(in, w, b, out, N, H, W, Cin, Cout, px_stride, st);
  else if (Cin % 8 == 0 && Cout % 8 == 0)
    launch_conv<8, 8, TIn>(in, w, b, out, N, H, W, Cin, Cout, px_stride, st);
  else
    return cudaErrorInvalidValue;
  return cudaGetLastError();
}
template cudaError_t Conv3x3Relu6F32<float>(const float*, const float*, const float*, float*, int, int, int, int,
                                            int, cudaStream_t, int);
template cudaError_t Conv3x3Relu6F32<uint8_t>(const uint8_t*, const float*, const float*, float*, int, int, int,
                                              int, int, cudaStream_t, int);

cudaError_t CropResizeU8(const uint8_t* src, int W, int cy, int cx, uint8_t* dst, int S, const int* taps, int area2x,
                         cudaStream_t st) {
  dim3 grid((S + 127) / 128, S);
  crop_resize_u8_kernel<<<grid, 128, 0, st>>>(src, W, cy, cx, dst, S, taps, area2x);
  return cudaGetLastError();
}

cudaError_t CropResizeBatchU8(const uint8_t* arena, const CropDesc* descs, int n, uint8_t* dst, int S, cudaStream_t st) {
  dim3 grid((S + 127) / 128, S, n);
  crop_resize_batch_kernel<<<grid, 128, 0, st>>>(arena, descs, dst, S);
  return cudaGetLastError();
}

cudaError_t Yuv420CropU8(const uint8_t* y, const uint8_t* u, const uint8_t* v, const YuvFrame& f, uint8_t* dst, int S,
                         cudaStream_t st) {
  dim3 grid((S + 127) / 128, S);
  yuv420_crop_kernel<<<grid, 128, 0, st>>>(y, u, v, f, dst, S);
  return cudaGetLastError();
}

cudaError_t AvgPoolF32(const float* in, float* out, int N, int H, int W, int C, int k, int s, cudaStream_t st) {
  int OH = (H - k) / s + 1, OW = (W - k) / s + 1;
  size_t total = static_cast<size_t>(N) * OH * OW * C;
  int blocks = static_cast<int>(std::min<size_t>((total + 255) / 256, static_cast<size_t>(SmCount()) * 16));
  avgpool_kernel<<<blocks, 256, 0, st>>>(in, out, N, H, W, C, k, s, OH, OW);
  return cudaGetLastError();
}

cudaError_t JoinF32(const float* p, const float* src, float* out, const float* A, const float* B, const float* C,
                    int N, int S, int SS, int Ch, cudaStream_t st) {
  size_t total = static_cast<size_t>(N) * S * S * Ch;
  int blocks = static_cast<int>(std::min<size_t>((total + 255) / 256, static_cast<size_t>(SmCount()) * 16));
  join_kernel<<<blocks, 256, 0, st>>>(p, src, out, A, B, C, N, S, SS, Ch);
  return cudaGetLastError();
}

cudaError_t DenseTailF32(const float* flat, int N, int flat_len, const DenseParams& dp, long long* top1, float* probs,
                         float* logits, float* pre_relu6, cudaStream_t st) {
  int warps_per_block = 4;
  int blocks = (N + warps_per_block - 1) / warps_per_block;
  dense_tail_kernel<<<blocks, warps_per_block * 32, 0, st>>>(flat, N, flat_len, dp, top1, probs, logits, pre_relu6);
  return cudaGetLastError();
}

}  // namespace rn
