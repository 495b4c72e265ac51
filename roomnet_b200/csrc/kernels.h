// Kernel launch wrappers shared by the engine.  All kernels are hand-written for
// sm_100a; nothing here dispatches to cuDNN/cuBLAS.
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <cstddef>
#include <cstdint>

#include "jpeg_host.h"

namespace rn {

struct DenseParams {
  const float* w[4];  // [in][out], BN of the producer folded in
  const float* b[4];
  int out[4];
};

// Number of SMs of the current device (148 on B200; cached per device) - grids are sized in multiples of it.
int SmCount();

// ---- fp32 CUDA-core kernels (kernels_f32.cu) --------------------------------
template <typename TIn>
cudaError_t Conv3x3Relu6F32(const TIn* in, const float* w, const float* b, float* out, int N, int H, int W, int Cin,
                            int Cout, cudaStream_t st, int px_stride = 0);
cudaError_t AvgPoolF32(const float* in, float* out, int N, int H, int W, int C, int k, int s, cudaStream_t st);
cudaError_t JoinF32(const float* p, const float* src, float* out, const float* A, const float* B, const float* C,
                    int N, int S, int SS, int Ch, cudaStream_t st);
cudaError_t DenseTailF32(const float* flat, int N, int flat_len, const DenseParams& dp, long long* top1, float* probs,
                         float* logits, float* pre_relu6, cudaStream_t st);

// Centre crop + cv2.resize-compatible (INTER_LINEAR, uint8 fixed point) bilinear resize; `taps` = 8*S int32 on the device.
cudaError_t CropResizeU8(const uint8_t* src, int W, int cy, int cx, uint8_t* dst, int S, const int* taps, int area2x,
                         cudaStream_t st);
// One image of a batched crop + resize: raw HxWx3 bytes at arena + offset, row pitch W pixels, crop origin (cy, cx),
// crop side `side`.
struct CropDesc {
  unsigned long long offset;
  int W, cy, cx, side;
};
// One YUV_420_888 camera frame (android.media.Image planes) and the rotation of the frame-to-crop transform.
struct YuvFrame {
  int width, height, y_row_stride, uv_row_stride, uv_pixel_stride, rotation;
};
cudaError_t Yuv420CropU8(const uint8_t* y, const uint8_t* u, const uint8_t* v, const YuvFrame& f, uint8_t* dst, int S,
                         cudaStream_t st);
cudaError_t CropResizeBatchU8(const uint8_t* arena, const CropDesc* descs, int n, uint8_t* dst, int S, cudaStream_t st);

// ---- JPEG front end, device half (kernels_jpeg.cu; host half: jpeg_host.h) ----
// One component of one image: quantised coefficients [hblocks][wblocks][64] (natural order) -> sample plane of
// wblocks*8 x hblocks*8 bytes.
struct JpegPlaneDesc {
  unsigned long long coef_offset;   // first coefficient (int16 units) in the coefficient arena
  unsigned long long plane_offset;  // first byte of the plane in the sample arena (8-byte aligned)
  int wblocks, hblocks;
  int quant_index;  // row of the uint16[64] quantisation-table array
  int pad;
};
// One image: its planes, the chroma geometry, and where the oriented BGR image goes in the raw-image arena (rows of
// out_pitch pixels, the arena offset a multiple of 256 bytes).
struct JpegImageDesc {
  unsigned long long plane[3];
  unsigned long long out_offset;
  int pitch[3];
  int width, height, ncomp;
  int hs, vs;    // luma sampling factors (1 or 2); chroma is 1x1
  int cdw, cdh;  // chroma width / height in samples
  int orientation;
  int out_pitch;  // row pitch of the oriented image in pixels (a multiple of four)
};
cudaError_t JpegIdct(const int16_t* coefs, const JpegPlaneDesc* planes, int n_planes, const uint16_t* quant,
                     uint8_t* samples, cudaStream_t st);
cudaError_t JpegColor(const uint8_t* samples, const JpegImageDesc* images, int n_images, uint8_t* raw, cudaStream_t st);

// ---- JPEG front end: Huffman decoding on the device (kernels_jpeg_huff.cu) ----
// One file = one interleaved scan.  Offsets are into the batch's arenas.
struct HuffFileDesc {
  unsigned long long stream_off;   // bytes into the stream arena (a multiple of kSubseqBytes)
  unsigned long long coef_off[3];  // int16 units into the coefficient arena, per component
  unsigned long long dc_off;       // int16 units into the DC-difference arena (one per block, scan order)
  unsigned sub_base, n_sub;        // this file's subsequences in the per-subsequence arrays
  unsigned total_blocks, seg_blocks;
  int bpm, mcus_x, ncomp, table_index;  // table_index: first of the file's six DevHuffTable
  int wblocks[3], hblocks[3], comp_h[3], comp_v[3];
  alignas(8) unsigned char blk_comp[8];
  unsigned char blk_hh[8], blk_vv[8];
};
struct HuffBlockDesc {  // one CUDA block = 256 consecutive subsequences of one file
  int file;
  unsigned first_sub;
};
struct HuffBatch {
  const HuffFileDesc* files;
  const HuffBlockDesc* blocks;
  const DevHuffTable* tables;
  const unsigned char* streams;
  const int* sub_seg;
  unsigned long long* state;
  unsigned long long* start_used;
  unsigned* nblk;
  unsigned* local_off;
  unsigned* block_sum;
  int* block_has_start;
  unsigned* carry;
  int16_t* coefs;    // zero-initialised by the caller
  int16_t* dcdiff;
  int* dc_part;      // scratch of the DC scan: 2 * (n_files * 3 * 8) ints
  int* file_error;   // zero-initialised by the caller; non-zero = damaged stream, decode that file on the host
  int* changed;      // device scratch
  int* h_changed;    // pinned host scratch
  int n_files, n_blocks;
  int* rounds_out;   // optional: re-synchronisation launches it took
};
// Enqueues the whole decode on `st` (synchronises the stream between re-synchronisation rounds).
cudaError_t HuffDecode(const HuffBatch& b, cudaStream_t st);

// ---- 16-bit tensor-core path (kernels_tc.cu) --------------------------------
// Activation layout between tensor-core layers ("chunked rows"):
//   T[n][y][cb][x][8]  16-bit elements, cb = channel / 8.  One (n, y, cb) plane row
//   is W*16 contiguous bytes, so a row segment is ONE 1-D TMA bulk copy and lands in
//   shared memory already in the UMMA no-swizzle K-major core-matrix order.
enum class HalfKind : int { kF16 = 0, kBF16 = 1 };

struct TcConvLayer {
  int cin, cout;          // logical channels of the conv
  int in_side;            // input spatial size (square)
  int pool_k, pool_s;     // 0/0 = no pooling
  int out_side;           // pooled output size
  const void* w_packed;   // device, packed by PackTcWeights
  size_t w_bytes;         // bytes per cout-part
  int cout_parts;         // cout is processed in `cout_parts` passes of cout/cout_parts channels
  const float* bias;      // device [cout], already divided by 6 (the kernel clips with saturate)
  int amode = 0;          // 2 = conv0 (uint8 image pre-expanded by PrepU8)
  // fused residual join in the epilogue (null = plain pooled output): out = A*pool + B*resize(join_src) + C
  bool split = false;              // fp32-class path: hi + lo activation planes, [Wh | Wl] weights (cin = physical channels)
  const void* join_src = nullptr;  // chunked tensor of the block's first pooled output
  int join_src_side = 0;
  const float* join_abc = nullptr;  // device [3][cout]
};

// Size in bytes of an activation tensor in chunked layout, incl. over-read slack.
size_t ChunkedBytes(int n, int side, int channels);

// Host-side weight packer: HWIO fp64 -> per-part shared-memory image of the conv_tc kernel.
// `scale` multiplies every weight before rounding (stored-activation scale bookkeeping, see engine.cu);
// `in_scale` (optional, [cin]) multiplies the weights of one input channel on top of that.
size_t PackTcWeights(const double* w_hwio, int cin, int cout, int cout_parts, HalfKind kind, double scale,
                     void* out_host, const double* in_scale = nullptr);
// split paths: per part [Wh | Wl] with wh = round16(scale * w), wl = round16(scale * w - wh); `cin` = LOGICAL
// input channels (a multiple of 16).  Returns the bytes of one part.
size_t PackTcWeightsSplit(const double* w_hwio, int cin, int cout, int cout_parts, HalfKind kind, double scale,
                          void* out_host);
// Round a value to the 16-bit storage type of the tensor-core path and return it as a double.
double RoundToHalfKind(double v, HalfKind kind);

// conv0 on the tensor cores: weights [3][3][3][8] (uint8-folded) split into fp16 hi + lo halves.
size_t PackTcConv0Weights(const double* w_hwio, HalfKind kind, double scale, void* out_host);
// uint8 NHWC image -> pixel-pair chunks (values / 256, exact) consumed by conv0's A descriptor.
cudaError_t PrepU8(const uint8_t* in, void* out, int N, int S, HalfKind kind, cudaStream_t st, int px_bytes = 3);

cudaError_t ConvTc(const TcConvLayer& L, const void* in, void* out, int N, HalfKind kind, cudaStream_t st);

// Residual block 2 (conv2d_2 -> conv2d_3 + join, reference network.py:183-203 / :227) as one kernel
// (kernels_block2.cu): `l1`, `l2` are the TcConvLayer records of the two layers as ConvTc would take them
// (l2.join_abc / join_src_side describe the join; the residual source is `in` itself), `in` = the block's first
// pooled tensor, `out` = the joined block output.  P2 (the tensor between the layers) never reaches HBM.
bool Block2FusedSupported(const TcConvLayer& l1, const TcConvLayer& l2);
cudaError_t Block2Fused(const TcConvLayer& l1, const TcConvLayer& l2, const void* in, void* out, int N, HalfKind kind,
                        cudaStream_t st);

// conv0 (3->8, CUDA cores, fp32 math) + ReLU6 + 3x3/1 avg-pool, writes chunked 16-bit.
template <typename TIn>
cudaError_t Conv0PoolH(const TIn* in, const float* w, const float* b, void* out, int N, int S, HalfKind kind,
                       cudaStream_t st);
// Fused tail (conv8 .. softmax) for small maps; dbg8/dbg9 (optional) receive the NHWC outputs of conv blocks 8/9.
bool TailFusedSupported(int s7, int channels);
cudaError_t TailFused(const void* p7, int N, int S7, float in_scale, const float* w8, const float* b8, const float* w9,
                      const float* b9, const float* ja, const float* jb, const float* jc, const DenseParams& dp,
                      int flat_len, HalfKind kind, long long* top1, float* probs, float* logits, float* dbg8,
                      float* dbg9, cudaStream_t st, bool split = false);
// chunked 16-bit -> NHWC fp32
cudaError_t F32ToSplitChunked(const float* in, void* out, int N, int S, int C, int Cpad, float scale, HalfKind kind,
                              cudaStream_t st);
cudaError_t ChunkedToF32(const void* in, float* out, int N, int S, int Ch, HalfKind kind, float scale, cudaStream_t st,
                         bool split = false);

}  // namespace rn
