#include "engine.h"

#include <cmath>
#include <cstdlib>
#include <cstring>

#include "roomnet.h"

namespace rn {

// NVTX ranges around the host-side phases (visible in Nsight Systems; a few nanoseconds when no tool is attached)
#include <nvtx3/nvToolsExt.h>
namespace {
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
};
}  // namespace

#define RN_CUDA(expr)                                                                                   \
  do {                                                                                                  \
    cudaError_t _e = (expr);                                                                            \
    if (_e != cudaSuccess) {                                                                            \
      err_ = std::string(#expr) + ": " + cudaGetErrorString(_e) + " (device " + std::to_string(device_) + ")"; \
      return _e;                                                                                        \
    }                                                                                                   \
  } while (0)

size_t InputBytesPerPixel(InputKind k) { return k == InputKind::kF32Rgb ? 12 : (k == InputKind::kArgb8888 ? 4 : 3); }

namespace {
size_t InputBytesPerImage(const NetShape& s, InputKind k) {
  return static_cast<size_t>(s.im_side) * s.im_side * InputBytesPerPixel(k);
}
}  // namespace

Replica::Replica(int device, const NetShape& shape, int precision, int max_batch, int flags)
    : device_(device), precision_(precision), max_batch_(max_batch), shape_(shape) {
  layerwise_ = (flags & RN_FLAG_LAYERWISE) != 0;
  flags_ = flags;
  half_kind_ = (precision == RN_PREC_BF16 || precision == RN_PREC_BF16X3) ? HalfKind::kBF16 : HalfKind::kF16;
  // RN_PREC_FP32_TC: conv0..conv7 as three-product split-fp16 tensor-core layers (hi + lo activations, [Wh | Wl]
  // weights, fp32 epilogues); the small tail (conv8, conv9, dense head) is fp32 as on the 16-bit path
  // RN_PREC_BF16X3: the same three-product scheme with bf16 halves (16 mantissa bits per operand): the BF16 path that
  // meets the 2e-2 budget, at a third of the single-product rate
  split_ = precision == RN_PREC_FP32_TC || precision == RN_PREC_BF16X3;
  first_f32_layer_ = precision == RN_PREC_FP32 ? 0 : 8;
}

Replica::~Replica() {
  cudaSetDevice(device_);
  cudaDeviceSynchronize();
  for (void* p : allocs_) cudaFree(p);
  if (d_raw_) cudaFree(d_raw_);
  if (d_coef_) cudaFree(d_coef_);
  if (d_samples_) cudaFree(d_samples_);
  if (d_jmeta_) cudaFree(d_jmeta_);
  for (uint8_t* p : d_huff_)
    if (p) cudaFree(p);
  for (cudaEvent_t ev : ev_huff_up_)
    if (ev) cudaEventDestroy(ev);
  if (h_huff_flags_) cudaFreeHost(h_huff_flags_);
  for (int16_t* p : h_coef_)
    if (p) cudaFreeHost(p);
  for (cudaEvent_t e : prof_events_) cudaEventDestroy(e);
  for (int i = 0; i < 2; ++i)
    if (h_in_[i]) cudaFreeHost(h_in_[i]);
  for (int i = 0; i < kSlots; ++i) {
    if (h_out_[i]) cudaFreeHost(h_out_[i]);
    if (ev_h2d_[i]) cudaEventDestroy(ev_h2d_[i]);
    if (ev_done_[i]) cudaEventDestroy(ev_done_[i]);
  }
  for (auto& set : sets_) {
    if (set.stream) cudaStreamDestroy(set.stream);
    if (set.ev_done) cudaEventDestroy(set.ev_done);
  }
  if (ev_fork_) cudaEventDestroy(ev_fork_);
  if (compute_) cudaStreamDestroy(compute_);
  if (copy_) cudaStreamDestroy(copy_);
}

cudaError_t Replica::Alloc(void** p, size_t bytes) {
  cudaError_t e = cudaMalloc(p, bytes ? bytes : 16);
  if (e == cudaSuccess) allocs_.push_back(*p);
  return e;
}

cudaError_t Replica::Init() {
  int count = 0;
  RN_CUDA(cudaGetDeviceCount(&count));
  if (device_ < 0 || device_ >= count) {
    err_ = "CUDA device " + std::to_string(device_) + " does not exist (" + std::to_string(count) + " visible)";
    return cudaErrorInvalidDevice;
  }
  RN_CUDA(cudaSetDevice(device_));
  cudaDeviceProp prop;
  RN_CUDA(cudaGetDeviceProperties(&prop, device_));
  if (prop.major != 10) {
    err_ = "device " + std::to_string(device_) + " is sm_" + std::to_string(prop.major * 10 + prop.minor) +
           "; libroomnet is built for sm_100a only (no fallback path)";
    return cudaErrorNoKernelImageForDevice;
  }
  RN_CUDA(cudaStreamCreateWithFlags(&compute_, cudaStreamNonBlocking));
  RN_CUDA(cudaStreamCreateWithFlags(&copy_, cudaStreamNonBlocking));
  for (int i = 0; i < kSlots; ++i) {
    RN_CUDA(cudaEventCreateWithFlags(&ev_h2d_[i], cudaEventDisableTiming));
    RN_CUDA(cudaEventCreateWithFlags(&ev_done_[i], cudaEventDisableTiming));
  }
  const size_t B = static_cast<size_t>(max_batch_);
  const int C = shape_.num_classes;
  RN_CUDA(cudaEventCreateWithFlags(&ev_fork_, cudaEventDisableTiming));
  for (int si = 0; si < 2; ++si) {
  cur_ = &sets_[si];
  RN_CUDA(cudaStreamCreateWithFlags(&cur_->stream, cudaStreamNonBlocking));
  RN_CUDA(cudaEventCreateWithFlags(&cur_->ev_done, cudaEventDisableTiming));
  size_t scratch = 0;
  for (int i = first_f32_layer_ == 0 ? 0 : first_f32_layer_; i < kNumConvs; ++i) {
    const ConvShape& cs = shape_.conv[i];
    scratch = std::max(scratch, static_cast<size_t>(cs.conv_side) * cs.conv_side * cs.cout);
  }
  if (split_)  // float feed: conv0 runs on the fp32 kernels and is split into hi + lo planes afterwards
    scratch = std::max(scratch, static_cast<size_t>(shape_.conv[0].conv_side) * shape_.conv[0].conv_side * shape_.conv[0].cout);
  RN_CUDA(Alloc(reinterpret_cast<void**>(&cur_->conv_scratch), scratch * B * sizeof(float)));
  for (int i = 0; i < kNumConvs; ++i) {
    const ConvShape& cs = shape_.conv[i];
    size_t elems = static_cast<size_t>(cs.out_side) * cs.out_side * cs.cout * B;
    bool f32_needed = i >= first_f32_layer_ || i == first_f32_layer_ - 1 || (split_ && i == 0);
    if (f32_needed) {
      RN_CUDA(Alloc(reinterpret_cast<void**>(&cur_->pooled[i]), elems * sizeof(float)));
      if (cs.join_src >= 0) RN_CUDA(Alloc(reinterpret_cast<void**>(&cur_->joined[i]), elems * sizeof(float)));
    }
    if (i < first_f32_layer_) {
      // split tensors carry hi and lo planes; conv0's eight channels travel padded to sixteen there (conv2d_1 then has
      // the even number of channel chunks per half that the K = 16 MMA steps need)
      size_t bytes = ChunkedBytes(max_batch_, cs.out_side, split_ ? 2 * std::max(cs.cout, 16) : cs.cout);
      RN_CUDA(Alloc(&cur_->act_h[i], bytes));
      RN_CUDA(cudaMemset(cur_->act_h[i], 0, bytes));
      if (cs.join_src >= 0) {
        RN_CUDA(Alloc(&cur_->join_h[i], bytes));
        RN_CUDA(cudaMemset(cur_->join_h[i], 0, bytes));
      }
    }
  }
  if (first_f32_layer_ > 0) RN_CUDA(Alloc(&cur_->in_h, ChunkedBytes(max_batch_, shape_.im_side, 8)));
  RN_CUDA(Alloc(reinterpret_cast<void**>(&cur_->d_pre), B * C * sizeof(float)));
  }
  cur_ = &sets_[0];
  const size_t in_bytes = InputBytesPerImage(shape_, InputKind::kF32Rgb) * B;
  for (int i = 0; i < 2; ++i) RN_CUDA(cudaMallocHost(&h_in_[i], in_bytes));
  for (int i = 0; i < kSlots; ++i) {
    RN_CUDA(Alloc(&d_in_[i], in_bytes));
    // results: the tail kernel stores top-1 / probabilities / logits (56 bytes per image) straight into this pinned,
    // device-mapped host buffer - no device-to-host copy operations on the compute stream
    RN_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&h_out_[i]), B * (sizeof(long long) + 2 * C * sizeof(float)),
                          cudaHostAllocMapped | cudaHostAllocPortable));
  }
  return cudaSuccess;
}

cudaError_t Replica::UploadF32(const std::vector<double>& v, float** dptr) {
  std::vector<float> f(v.size());
  for (size_t i = 0; i < v.size(); ++i) f[i] = static_cast<float>(v[i]);
  if (!*dptr) RN_CUDA(Alloc(reinterpret_cast<void**>(dptr), f.size() * sizeof(float)));
  RN_CUDA(cudaMemcpy(*dptr, f.data(), f.size() * sizeof(float), cudaMemcpyHostToDevice));
  return cudaSuccess;
}

cudaError_t Replica::Upload(const FoldedNet& f) {
  NvtxRange nvtx_range("rn::Upload (pack + upload folded weights)");
  RN_CUDA(cudaSetDevice(device_));
  RN_CUDA(cudaDeviceSynchronize());
  const FoldedConv* c0[3] = {&f.conv0_u8bgr, &f.conv0_u8rgb, &f.conv0_f32rgb};
  for (int k = 0; k < 3; ++k) {
    RN_CUDA(UploadF32(c0[k]->w, &w0_[k]));
    RN_CUDA(UploadF32(c0[k]->b, &b0_[k]));
  }
  for (int i = 1; i < kNumConvs; ++i) {
    RN_CUDA(UploadF32(f.conv[i].w, &cw_[i]));
    RN_CUDA(UploadF32(f.conv[i].b, &cb_[i]));
  }
  for (int i = 0; i < kNumConvs; ++i) {
    if (shape_.conv[i].join_src < 0) continue;
    RN_CUDA(UploadF32(f.join[i].a, &ja_[i]));
    RN_CUDA(UploadF32(f.join[i].b, &jb_[i]));
    RN_CUDA(UploadF32(f.join[i].c, &jc_[i]));
  }
  float* dw[4] = {const_cast<float*>(dense_.w[0]), const_cast<float*>(dense_.w[1]), const_cast<float*>(dense_.w[2]),
                  const_cast<float*>(dense_.w[3])};
  float* db[4] = {const_cast<float*>(dense_.b[0]), const_cast<float*>(dense_.b[1]), const_cast<float*>(dense_.b[2]),
                  const_cast<float*>(dense_.b[3])};
  for (int i = 0; i < kNumDense; ++i) {
    RN_CUDA(UploadF32(f.dense[i].w, &dw[i]));
    RN_CUDA(UploadF32(f.dense[i].b, &db[i]));
    dense_.w[i] = dw[i];
    dense_.b[i] = db[i];
    dense_.out[i] = shape_.dense_out[i];
  }
  if (precision_ != RN_PREC_FP32) {
    // conv0: both the tensor-core kernel (uint8 feeds) and the CUDA-core kernel (float feed) store
    // sum_3x3(relu6)/6 = 1.5 x the pooled mean
    act_scale_[0] = 9.0 / 6.0;
    {
      const ConvShape& c0 = shape_.conv[0];
      TcConvLayer& L = tc_[0];
      L.cin = 8;
      L.cout = 16;
      L.in_side = c0.in_side;
      L.pool_k = c0.pool_k;
      L.pool_s = c0.pool_s;
      L.out_side = c0.out_side;
      L.cout_parts = 1;
      L.amode = 2;
      L.split = split_;
      std::vector<double> b6(16, 0.0);
      for (int k = 0; k < 8; ++k) b6[k] = f.conv0_u8bgr.b[k] / 6.0;  // identical for BGR and RGB byte order
      RN_CUDA(UploadF32(b6, &tc_bias_[0]));
      L.bias = tc_bias_[0];
      L.w_bytes = PackTcConv0Weights(nullptr, half_kind_, 1.0, nullptr);
      const FoldedConv* src[2] = {&f.conv0_u8bgr, &f.conv0_u8rgb};
      for (int k = 0; k < 2; ++k) {
        std::vector<uint8_t> host(L.w_bytes);
        // activations are fed as pixel/256 (exact), so the weights carry 256/6
        PackTcConv0Weights(src[k]->w.data(), half_kind_, 256.0 / 6.0, host.data());
        if (!tc0_w_[k]) RN_CUDA(Alloc(&tc0_w_[k], host.size()));
        RN_CUDA(cudaMemcpy(tc0_w_[k], host.data(), host.size(), cudaMemcpyHostToDevice));
      }
    }
    for (int i = 1; i < first_f32_layer_; ++i) {
      const ConvShape& cs = shape_.conv[i];
      const bool in_is_join = shape_.conv[i - 1].join_src >= 0;
      const double g_in = in_is_join ? 1.0 : act_scale_[i - 1];
      // the kernel computes saturate(conv(stored_in, W') + b') == relu6(z)/6 and stores the pool-window SUM
      act_scale_[i] = (cs.pool_k ? cs.pool_k * cs.pool_k : 1) / 6.0;
      TcConvLayer& L = tc_[i];
      const int cin_l = split_ ? std::max(cs.cin, 16) : cs.cin;  // logical input channels (conv2d_1: 8 real + 8 zero)
      L.split = split_;
      L.cin = split_ ? 2 * cin_l : cs.cin;
      L.cout = cs.cout;
      L.in_side = cs.in_side;
      L.pool_k = cs.pool_k;
      L.pool_s = cs.pool_s;
      L.out_side = cs.out_side;
      // split layers keep [Wh | Wl] and hi + lo stages in shared memory: 64-channel inputs leave room for 32 outputs
      L.cout_parts = split_ ? (cin_l >= 64 ? std::max(1, cs.cout / 32) : 1) : (cs.cout > 64 ? cs.cout / 64 : 1);
      std::vector<double> b6(f.conv[i].b.size());
      for (size_t k = 0; k < b6.size(); ++k) b6[k] = f.conv[i].b[k] / 6.0;
      RN_CUDA(UploadF32(b6, &tc_bias_[i]));
      L.bias = tc_bias_[i];
      size_t part_bytes = 0;
      std::vector<uint8_t> host;
      if (split_) {
        std::vector<double> wpad(static_cast<size_t>(9) * cin_l * cs.cout, 0.0);
        for (int t = 0; t < 9; ++t)
          for (int c = 0; c < cs.cin; ++c)
            for (int o = 0; o < cs.cout; ++o)
              wpad[(static_cast<size_t>(t) * cin_l + c) * cs.cout + o] = f.conv[i].w[(static_cast<size_t>(t) * cs.cin + c) * cs.cout + o];
        part_bytes = PackTcWeightsSplit(nullptr, cin_l, cs.cout, L.cout_parts, half_kind_, 1.0, nullptr);
        host.resize(part_bytes * L.cout_parts);
        PackTcWeightsSplit(wpad.data(), cin_l, cs.cout, L.cout_parts, half_kind_, 1.0 / (6.0 * g_in), host.data());
      } else {
      part_bytes = PackTcWeights(nullptr, cs.cin, cs.cout, L.cout_parts, half_kind_, 1.0, nullptr);
      host.resize(part_bytes * L.cout_parts);
      std::vector<double> in_scale;  // the producer's per-channel join gain (below) is undone in this layer's weights
      if (in_is_join && !join_gain_[i - 1].empty())
        for (double g : join_gain_[i - 1]) in_scale.push_back(1.0 / g);
      PackTcWeights(f.conv[i].w.data(), cs.cin, cs.cout, L.cout_parts, half_kind_, 1.0 / (6.0 * g_in), host.data(),
                    in_scale.empty() ? nullptr : in_scale.data());
      }
      void* d = const_cast<void*>(L.w_packed);
      if (!d) RN_CUDA(Alloc(&d, host.size()));
      RN_CUDA(cudaMemcpy(d, host.data(), host.size(), cudaMemcpyHostToDevice));
      L.w_packed = d;
      L.w_bytes = part_bytes;
      if (cs.join_src >= 0) {  // out = (A/g_k) * stored_k + (B/g_0) * resize(stored_0) + C  -> true scale
        std::vector<double> a(f.join[i].a), b(f.join[i].b);
        for (auto& v : a) v /= act_scale_[i];
        for (auto& v : b) v /= act_scale_[cs.join_src];
        std::vector<double> cc(f.join[i].c);
        join_gain_[i].clear();
        if (i + 1 < first_f32_layer_ && !split_) {  // (split layers join in fp32: no gain needed)
          // The tensor-core kernels (kernels_block2.cu, the JOIN epilogue of kernels_tc.cu) multiply the resized
          // residual by B with a mixed-precision fma whose multiplier is a 16-bit value.  Rounding B would be a coherent per-channel error of 2^-12 (measured: 1e-2 on
          // the logits of flat images), so the whole output channel is stored with a gain g = round16(B) / B instead:
          // g*A and g*C stay fp32, g*B is exactly representable, and the next conv's weights of that input channel carry 1/g.
          join_gain_[i].assign(cs.cout, 1.0);
          for (int k = 0; k < cs.cout; ++k) {
            const double bh = RoundToHalfKind(b[k], half_kind_);
            if (b[k] != 0.0 && bh != 0.0) {
              join_gain_[i][k] = bh / b[k];
              a[k] *= join_gain_[i][k];
              cc[k] *= join_gain_[i][k];
              b[k] = bh;
            }
          }
        }
        std::vector<double> abc(a);
        abc.insert(abc.end(), b.begin(), b.end());
        abc.insert(abc.end(), cc.begin(), cc.end());
        RN_CUDA(UploadF32(abc, &tc_abc_[i]));
        L.join_abc = tc_abc_[i];
        L.join_src_side = shape_.conv[cs.join_src].out_side;
      }
    }
  }
  loaded_ = true;
  return cudaSuccess;
}

void Replica::Mark(const char* name, cudaStream_t st) {
  if (name) ++last_launches_;
  if (!profiling_) return;
  if (prof_used_ == prof_events_.size()) {
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) return;
    prof_events_.push_back(e);
    prof_names_.push_back("");
  }
  cudaEventRecord(prof_events_[prof_used_], st);
  prof_names_[prof_used_] = name ? name : "";  // name of the kernel that ENDS at this event
  ++prof_used_;
}

cudaError_t Replica::ProfileResults(std::vector<KernelTime>* out) {
  RN_CUDA(cudaSetDevice(device_));
  RN_CUDA(cudaDeviceSynchronize());
  out->clear();
  for (size_t i = 1; i < prof_used_; ++i) {
    if (prof_names_[i].empty()) continue;  // start marker of a micro-batch
    float ms = 0.f;
    RN_CUDA(cudaEventElapsedTime(&ms, prof_events_[i - 1], prof_events_[i]));
    auto it = std::find_if(out->begin(), out->end(), [&](const KernelTime& k) { return k.name == prof_names_[i]; });
    if (it == out->end()) {
      out->push_back(KernelTime{prof_names_[i], 0.0, 0});
      it = out->end() - 1;
    }
    it->ms += ms;
    it->launches += 1;
  }
  prof_used_ = 0;
  return cudaSuccess;
}

cudaError_t Replica::TailF32(int first_layer, int n, cudaStream_t st) {
  for (int i = first_layer; i < kNumConvs; ++i) {
    const ConvShape& cs = shape_.conv[i];
    const float* in = shape_.conv[i - 1].join_src >= 0 ? cur_->joined[i - 1] : cur_->pooled[i - 1];
    float* conv_out = cs.pool_k ? cur_->conv_scratch : cur_->pooled[i];
    RN_CUDA(Conv3x3Relu6F32<float>(in, cw_[i], cb_[i], conv_out, n, cs.in_side, cs.in_side, cs.cin, cs.cout, st));
    Mark(("conv" + std::to_string(i) + "_f32").c_str(), st);
    if (cs.pool_k) {
      RN_CUDA(AvgPoolF32(cur_->conv_scratch, cur_->pooled[i], n, cs.conv_side, cs.conv_side, cs.cout, cs.pool_k, cs.pool_s, st));
      Mark(("pool" + std::to_string(i) + "_f32").c_str(), st);
    }
    if (cs.join_src >= 0) {
      RN_CUDA(JoinF32(cur_->pooled[i], cur_->pooled[cs.join_src], cur_->joined[i], ja_[i], jb_[i], jc_[i], n, cs.out_side,
                      shape_.conv[cs.join_src].out_side, cs.cout, st));
      Mark(("join" + std::to_string(i) + "_f32").c_str(), st);
    }
  }
  return cudaSuccess;
}

cudaError_t Replica::ForwardF32(const void* d_in, InputKind kind, int n, cudaStream_t st) {
  const ConvShape& c0 = shape_.conv[0];
  const bool argb = kind == InputKind::kArgb8888;  // BGRA bytes: the BGR weights with a 4-byte pixel pitch
  const int k = static_cast<int>(argb ? InputKind::kU8Bgr : kind);
  Mark(nullptr, st);
  if (kind == InputKind::kF32Rgb)
    RN_CUDA(Conv3x3Relu6F32<float>(static_cast<const float*>(d_in), w0_[k], b0_[k], cur_->conv_scratch, n, c0.in_side,
                                   c0.in_side, 3, c0.cout, st));
  else
    RN_CUDA(Conv3x3Relu6F32<uint8_t>(static_cast<const uint8_t*>(d_in), w0_[k], b0_[k], cur_->conv_scratch, n, c0.in_side,
                                     c0.in_side, 3, c0.cout, st, argb ? 4 : 3));
  Mark("conv0_f32", st);
  RN_CUDA(AvgPoolF32(cur_->conv_scratch, cur_->pooled[0], n, c0.conv_side, c0.conv_side, c0.cout, c0.pool_k, c0.pool_s, st));
  Mark("pool0_f32", st);
  return TailF32(1, n, st);
}

cudaError_t Replica::ForwardTc(const void* d_in, InputKind kind, int n, long long* d_top1, float* d_probs,
                               float* d_logits, cudaStream_t st) {
  const ConvShape& c0 = shape_.conv[0];
  const bool argb = kind == InputKind::kArgb8888;
  const int k = static_cast<int>(argb ? InputKind::kU8Bgr : kind);
  Mark(nullptr, st);
  if (kind == InputKind::kF32Rgb && split_) {
    // raw float feed on the fp32-class path: conv0 + pool on the fp32 kernels, then the split into hi + lo planes
    RN_CUDA(Conv3x3Relu6F32<float>(static_cast<const float*>(d_in), w0_[k], b0_[k], cur_->conv_scratch, n, c0.in_side,
                                   c0.in_side, 3, c0.cout, st));
    Mark("conv0_f32", st);
    RN_CUDA(AvgPoolF32(cur_->conv_scratch, cur_->pooled[0], n, c0.conv_side, c0.conv_side, c0.cout, c0.pool_k, c0.pool_s, st));
    Mark("pool0_f32", st);
    RN_CUDA(F32ToSplitChunked(cur_->pooled[0], cur_->act_h[0], n, c0.out_side, c0.cout, 16,
                              static_cast<float>(act_scale_[0]), half_kind_, st));
    Mark("f32_to_split", st);
  } else if (kind == InputKind::kF32Rgb) {
    // raw float feed: operands need more than 11 bits, keep conv0 in fp32 on the CUDA cores
    RN_CUDA(Conv0PoolH<float>(static_cast<const float*>(d_in), w0_[k], b0_[k], cur_->act_h[0], n, c0.in_side, half_kind_, st));
    Mark("conv0_pool_h", st);
  } else {
    RN_CUDA(PrepU8(static_cast<const uint8_t*>(d_in), cur_->in_h, n, c0.in_side, half_kind_, st, argb ? 4 : 3));
    Mark("prep_u8", st);
    TcConvLayer L0 = tc_[0];
    L0.w_packed = tc0_w_[k];
    RN_CUDA(ConvTc(L0, cur_->in_h, cur_->act_h[0], n, half_kind_, st));
    Mark("conv0_tc", st);
  }
  for (int i = 1; i < first_f32_layer_; ++i) {
    const ConvShape& cs = shape_.conv[i];
    const void* in = shape_.conv[i - 1].join_src >= 0 ? cur_->join_h[i - 1] : cur_->act_h[i - 1];
    if (i == 2 && !layerwise_ && !split_ && shape_.conv[3].join_src == 1 && Block2FusedSupported(tc_[2], tc_[3])) {
      // residual block 2 in one kernel: conv2d_2's output stays in shared memory (kernels_block2.cu)
      RN_CUDA(Block2Fused(tc_[2], tc_[3], cur_->act_h[1], cur_->join_h[3], n, half_kind_, st));
      Mark("block2_tc", st);
      block2_fused_last_ = true;
      i = 3;
      continue;
    }
    if (cs.join_src >= 0) {  // the residual join runs in the conv epilogue
      TcConvLayer L = tc_[i];
      L.join_src = cur_->act_h[cs.join_src];
      RN_CUDA(ConvTc(L, in, cur_->join_h[i], n, half_kind_, st));
    } else {
      RN_CUDA(ConvTc(tc_[i], in, cur_->act_h[i], n, half_kind_, st));
    }
    Mark(("conv" + std::to_string(i) + "_tc").c_str(), st);
  }
  const int last = first_f32_layer_ - 1;
  const ConvShape& cl = shape_.conv[last];
  if (last == 7 && TailFusedSupported(cl.out_side, cl.cout) && shape_.conv[8].pool_k == 4 && shape_.conv[8].pool_s == 2 &&
      shape_.conv[9].pool_k == 4 && shape_.conv[9].pool_s == 2 && shape_.conv[9].join_src == 7) {
    RN_CUDA(TailFused(cur_->act_h[7], n, cl.out_side, static_cast<float>(1.0 / act_scale_[7]), cw_[8], cb_[8], cw_[9], cb_[9],
                      ja_[9], jb_[9], jc_[9], dense_, shape_.flat_len, half_kind_, d_top1, d_probs, d_logits, cur_->pooled[8],
                      cur_->joined[9], st, split_));
    Mark("tail_fused", st);
    return cudaSuccess;
  }
  RN_CUDA(ChunkedToF32(cur_->act_h[last], cur_->pooled[last], n, cl.out_side, cl.cout, half_kind_,
                       static_cast<float>(1.0 / act_scale_[last]), st, split_));
  Mark("chunked_to_f32", st);
  cudaError_t e = TailF32(first_f32_layer_, n, st);
  if (e != cudaSuccess) return e;
  return DenseTail(n, d_top1, d_probs, d_logits, st);
}

cudaError_t Replica::DenseTail(int n, long long* d_top1, float* d_probs, float* d_logits, cudaStream_t st) {
  const ConvShape& cl = shape_.conv[kNumConvs - 1];
  const float* flat = cl.join_src >= 0 ? cur_->joined[kNumConvs - 1] : cur_->pooled[kNumConvs - 1];
  RN_CUDA(DenseTailF32(flat, n, shape_.flat_len, dense_, d_top1, d_probs, d_logits, cur_->d_pre, st));
  Mark("dense_tail", st);
  return cudaSuccess;
}

cudaError_t Replica::ForwardDevice(const void* d_in, InputKind kind, int n, long long* d_top1, float* d_probs,
                                   float* d_logits, cudaStream_t st) {
  if (n <= 0 || n > max_batch_) {
    err_ = "micro-batch size out of range";
    return cudaErrorInvalidValue;
  }
  cudaError_t e;
  if (precision_ == RN_PREC_FP32) {
    e = ForwardF32(d_in, kind, n, st);
    if (e == cudaSuccess) e = DenseTail(n, d_top1, d_probs, d_logits, st);
  } else {
    e = ForwardTc(d_in, kind, n, d_top1, d_probs, d_logits, st);
  }
  if (e != cudaSuccess) return e;
  cur_->last_n = n;
  return cudaSuccess;
}

cudaError_t Replica::InferDevice(const void* d_in, InputKind kind, int n, long long* d_top1, float* d_probs,
                                 float* d_logits, cudaStream_t st) {
  NvtxRange nvtx_range("rn::InferDevice");
  RN_CUDA(cudaSetDevice(device_));
  if (!st) st = compute_;
  last_launches_ = 0;
  const size_t per = InputBytesPerImage(shape_, kind);
  const int C = shape_.num_classes;
  // Two half-size micro-batches on two streams overlap better than one big one (see ActSet); while
  // profiling everything stays on `st` so that the per-kernel events measure isolated kernels.
  const bool overlap = !profiling_ && n >= 64;
  const int chunk = overlap ? std::min(max_batch_, std::max(32, (n + 1) / 2)) : max_batch_;
  // the activation sets are shared with the host entry points, whose work runs on the sets' own streams: order this
  // call after whatever they still have in flight (and, below, later host calls after this one)
  for (int si = 0; si < 2; ++si) {
    RN_CUDA(cudaEventRecord(sets_[si].ev_done, sets_[si].stream));
    RN_CUDA(cudaStreamWaitEvent(st, sets_[si].ev_done, 0));
  }
  if (overlap) RN_CUDA(cudaEventRecord(ev_fork_, st));
  int k = 0;
  for (int off = 0; off < n; off += chunk, ++k) {
    const int m = std::min(chunk, n - off);
    cur_ = &sets_[overlap ? (k & 1) : 0];
    cudaStream_t s = overlap ? cur_->stream : st;
    if (overlap && k < 2) RN_CUDA(cudaStreamWaitEvent(s, ev_fork_, 0));
    cudaError_t e = ForwardDevice(static_cast<const char*>(d_in) + per * off, kind, m, d_top1 ? d_top1 + off : nullptr,
                                  d_probs ? d_probs + static_cast<size_t>(off) * C : nullptr,
                                  d_logits ? d_logits + static_cast<size_t>(off) * C : nullptr, s);
    if (e != cudaSuccess) return e;
  }
  if (overlap) {
    for (int si = 0; si < std::min(k, 2); ++si) {
      RN_CUDA(cudaEventRecord(sets_[si].ev_done, sets_[si].stream));
      RN_CUDA(cudaStreamWaitEvent(st, sets_[si].ev_done, 0));
    }
  } else {
    RN_CUDA(cudaEventRecord(ev_fork_, st));
    RN_CUDA(cudaStreamWaitEvent(sets_[0].stream, ev_fork_, 0));
  }
  return cudaSuccess;
}

cudaError_t Replica::DrainSlot(int slot) {
  PendingOut& q = pend_[slot];
  if (!q.active) return cudaSuccess;
  q.active = false;
  RN_CUDA(cudaEventSynchronize(ev_done_[slot]));
  Deliver(slot, q.m, q.top1, q.probs, q.logits);
  return cudaSuccess;
}

Replica::HostOut Replica::Out(int slot) const {
  const int C = shape_.num_classes;
  char* base = h_out_[slot];
  return HostOut{reinterpret_cast<long long*>(base), reinterpret_cast<float*>(base + max_batch_ * sizeof(long long)),
                 reinterpret_cast<float*>(base + max_batch_ * (sizeof(long long) + C * sizeof(float)))};
}

void Replica::Deliver(int slot, int m, int64_t* top1, float* probs, float* logits) const {
  const int C = shape_.num_classes;
  const HostOut ho = Out(slot);
  if (top1) std::memcpy(top1, ho.top1, m * sizeof(long long));
  if (probs) std::memcpy(probs, ho.probs, static_cast<size_t>(m) * C * sizeof(float));
  if (logits) std::memcpy(logits, ho.logits, static_cast<size_t>(m) * C * sizeof(float));
}

// error path: nothing of this replica may still be running (or be delivered) when the failing call returns
void Replica::AbortPending() {
  cudaDeviceSynchronize();
  for (auto& q : pend_) q.active = false;
}

cudaError_t Replica::WaitHost(uint64_t ticket) {
  NvtxRange nvtx_range("rn::WaitHost (deliver results)");
  RN_CUDA(cudaSetDevice(device_));
  for (int s = 0; s < kSlots; ++s) {  // oldest slot first
    const int slot = (slot_seq_ + s) % kSlots;
    if (pend_[slot].active && pend_[slot].ticket <= ticket) {
      cudaError_t e = DrainSlot(slot);
      if (e != cudaSuccess) {
        AbortPending();
        return e;
      }
    }
  }
  return cudaSuccess;
}

cudaError_t Replica::InferHost(const void* h_in, InputKind kind, int n, int64_t* top1, float* probs, float* logits) {
  cudaError_t e = WaitHost(~0ull);  // a synchronous call comes after everything submitted before it
  if (e != cudaSuccess) return e;
  e = SubmitHost(h_in, kind, n, top1, probs, logits, 0);
  if (e != cudaSuccess) return e;
  return WaitHost(~0ull);
}

cudaError_t Replica::SubmitHost(const void* h_in, InputKind kind, int n, int64_t* top1, float* probs, float* logits,
                                uint64_t ticket) {
  NvtxRange nvtx_range("rn::SubmitHost (copies + kernels of one call)");
  RN_CUDA(cudaSetDevice(device_));
  last_launches_ = 0;
  const size_t per = InputBytesPerImage(shape_, kind);
  const int C = shape_.num_classes;
  cudaPointerAttributes attr;
  bool pinned_in = cudaPointerGetAttributes(&attr, h_in) == cudaSuccess && attr.type == cudaMemoryTypeHost;
  cudaGetLastError();  // cudaPointerGetAttributes on pageable memory may set a sticky-less error
  // Micro-batch schedule of one call: the first H2D copy is exposed when nothing is in flight, so the first
  // micro-batch is small; the rest are as large as possible (kernel efficiency) while still alternating between
  // the two activation sets so that copies overlap the previous micro-batches' kernels.
  std::vector<int> sizes;
  bool idle = true;
  for (const auto& q : pend_) idle = idle && !q.active;
  if (n >= 128) {
    int rest = n;
    if (idle) {
      const int first = std::min(max_batch_, std::max(32, n / 4));
      sizes.push_back(first);
      rest -= first;
    }
    const int parts = std::max(2, (rest + max_batch_ - 1) / max_batch_);
    for (int i = 0; i < parts; ++i) {
      const int m = rest / (parts - i);
      if (m > 0) sizes.push_back(m);
      rest -= m;
    }
  } else {
    for (int off = 0; off < n; off += max_batch_) sizes.push_back(std::min(max_batch_, n - off));
  }
  auto fail = [&](cudaError_t e) {
    err_ = std::string("SubmitHost: ") + cudaGetErrorString(e) + " (device " + std::to_string(device_) + ")";
    AbortPending();
    return e;
  };
  int off = 0;
  for (size_t si = 0; si < sizes.size(); off += sizes[si], ++si, ++slot_seq_) {
    const int slot = slot_seq_ % kSlots;
    const int m = sizes[si];
    cudaError_t e = DrainSlot(slot);  // the slot's device / host buffers are free after this
    if (e != cudaSuccess) return fail(e);
    const char* src = static_cast<const char*>(h_in) + per * off;
    if (!pinned_in) {  // pageable caller memory: through a pinned bounce buffer, reusable once its last copy is done
      const int hb = slot_seq_ & 1;
      if (slot_seq_ >= 2 && (e = cudaEventSynchronize(ev_h2d_[(slot_seq_ - 2) % kSlots])) != cudaSuccess) return fail(e);
      std::memcpy(h_in_[hb], src, per * m);
      src = static_cast<const char*>(h_in_[hb]);
    }
    if ((e = cudaMemcpyAsync(d_in_[slot], src, per * m, cudaMemcpyHostToDevice, copy_)) != cudaSuccess) return fail(e);
    if ((e = cudaEventRecord(ev_h2d_[slot], copy_)) != cudaSuccess) return fail(e);
    cur_ = &sets_[profiling_ ? 0 : (slot_seq_ & 1)];
    cudaStream_t cs = profiling_ ? compute_ : cur_->stream;
    if ((e = cudaStreamWaitEvent(cs, ev_h2d_[slot], 0)) != cudaSuccess) return fail(e);
    const HostOut ho = Out(slot);
    e = ForwardDevice(d_in_[slot], kind, m, ho.top1, ho.probs, ho.logits, cs);
    if (e != cudaSuccess) {
      AbortPending();
      return e;
    }
    if ((e = cudaEventRecord(ev_done_[slot], cs)) != cudaSuccess) return fail(e);
    PendingOut& q = pend_[slot];
    q.m = m;
    q.active = true;
    q.ticket = ticket;
    q.top1 = top1 ? top1 + off : nullptr;
    q.probs = probs ? probs + static_cast<size_t>(off) * C : nullptr;
    q.logits = logits ? logits + static_cast<size_t>(off) * C : nullptr;
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(e);
  return cudaSuccess;
}

namespace {
// One axis of OpenCV's INTER_LINEAR tap table for uint8 (resize.cpp): index pair + 11-bit weights.
void ResizeTaps(int dst, int src, bool vertical, int* s0, int* s1, int* w0, int* w1) {
  const double scale = static_cast<double>(src) / static_cast<double>(dst);
  for (int d = 0; d < dst; ++d) {
    float f = static_cast<float>((d + 0.5) * scale - 0.5);
    int s = static_cast<int>(std::floor(f));
    f -= static_cast<float>(s);
    int a, b;
    if (vertical) {  // rows: the fraction is kept, only the two row indices are clipped
      a = std::min(std::max(s, 0), src - 1);
      b = std::min(std::max(s + 1, 0), src - 1);
    } else {         // columns: out-of-range taps fold by clamping the fraction
      if (s < 0) {
        f = 0.f;
        s = 0;
      }
      if (s >= src - 1) {
        f = 0.f;
        s = src - 1;
      }
      a = s;
      b = std::min(s + 1, src - 1);
    }
    s0[d] = a;
    s1[d] = b;
    w1[d] = static_cast<int>(std::nearbyint(f * 2048.0f));
    w0[d] = static_cast<int>(std::nearbyint((1.0f - f) * 2048.0f));
  }
}
}  // namespace

cudaError_t Replica::Preprocess(const uint8_t* h_img, int H, int W, uint8_t* h_out) {
  {
    cudaError_t ew = WaitHost(~0ull);  // staging slot 0 is used below: nothing submitted earlier may still own it
    if (ew != cudaSuccess) return ew;
  }
  RN_CUDA(cudaSetDevice(device_));
  const int S = shape_.im_side;
  // reference network.py:139: offset = abs((w - h) // 2) with Python floor division
  const int d = W - H;
  const int fl = d >= 0 ? d / 2 : -((-d + 1) / 2);
  const int off = fl < 0 ? -fl : fl;
  const int side = std::min(H, W), cy = H > W ? off : 0, cx = W > H ? off : 0;
  const size_t raw_bytes = static_cast<size_t>(H) * W * 3;
  if (raw_bytes > d_raw_cap_) {
    if (d_raw_) cudaFree(d_raw_);
    d_raw_ = nullptr;
    d_raw_cap_ = 0;
    RN_CUDA(cudaMalloc(reinterpret_cast<void**>(&d_raw_), raw_bytes));
    d_raw_cap_ = raw_bytes;
  }
  if (!d_taps_) RN_CUDA(Alloc(reinterpret_cast<void**>(&d_taps_), static_cast<size_t>(8) * S * sizeof(int)));
  RN_CUDA(cudaMemcpyAsync(d_raw_, h_img, raw_bytes, cudaMemcpyHostToDevice, compute_));
  uint8_t* dst = static_cast<uint8_t*>(d_in_[0]);
  if (side == S) {  // already the right size: plain crop copy (network.py:151 skips the resize)
    RN_CUDA(cudaMemcpy2DAsync(dst, static_cast<size_t>(S) * 3, d_raw_ + (static_cast<size_t>(cy) * W + cx) * 3,
                              static_cast<size_t>(W) * 3, static_cast<size_t>(S) * 3, S, cudaMemcpyDeviceToDevice,
                              compute_));
  } else {
    std::vector<int> taps(static_cast<size_t>(8) * S);
    ResizeTaps(S, side, false, &taps[0], &taps[S], &taps[2 * S], &taps[3 * S]);
    ResizeTaps(S, side, true, &taps[4 * S], &taps[5 * S], &taps[6 * S], &taps[7 * S]);
    RN_CUDA(cudaMemcpyAsync(d_taps_, taps.data(), taps.size() * sizeof(int), cudaMemcpyHostToDevice, compute_));
    RN_CUDA(cudaStreamSynchronize(compute_));  // `taps` is a pageable temporary
    RN_CUDA(CropResizeU8(d_raw_, W, cy, cx, dst, S, d_taps_, side == 2 * S ? 1 : 0, compute_));
  }
  if (h_out) RN_CUDA(cudaMemcpyAsync(h_out, dst, static_cast<size_t>(S) * S * 3, cudaMemcpyDeviceToHost, compute_));
  RN_CUDA(cudaStreamSynchronize(compute_));
  return cudaSuccess;
}

cudaError_t Replica::InferImages(const uint8_t* const* imgs, const int* H, const int* W, int n, int64_t* top1,
                                 float* probs, float* logits) {
  NvtxRange nvtx_range("rn::InferImages (batched crop + resize + forward)");
  {
    cudaError_t ew = WaitHost(~0ull);  // uses staging slot 0 and activation set 0 on the replica's own stream
    if (ew != cudaSuccess) return ew;
  }
  RN_CUDA(cudaSetDevice(device_));
  last_launches_ = 0;
  const int S = shape_.im_side, C = shape_.num_classes;
  if (!d_descs_) RN_CUDA(Alloc(&d_descs_, static_cast<size_t>(max_batch_) * sizeof(CropDesc)));
  std::vector<CropDesc> descs;
  for (int off = 0; off < n; off += max_batch_) {
    const int m = std::min(max_batch_, n - off);
    descs.assign(m, CropDesc{});
    size_t total = 0;
    for (int i = 0; i < m; ++i) {
      const int h = H[off + i], w = W[off + i];
      if (!imgs[off + i] || h <= 0 || w <= 0) {
        err_ = "image " + std::to_string(off + i) + ": null pointer or empty size";
        return cudaErrorInvalidValue;
      }
      // reference network.py:139: offset = abs((w - h) // 2) with Python floor division
      const int d = w - h;
      const int fl = d >= 0 ? d / 2 : -((-d + 1) / 2);
      const int o = fl < 0 ? -fl : fl;
      descs[i].offset = total;
      descs[i].W = w;
      descs[i].side = std::min(h, w);
      descs[i].cy = h > w ? o : 0;
      descs[i].cx = w > h ? o : 0;
      total += (static_cast<size_t>(h) * w * 3 + 255) & ~static_cast<size_t>(255);
    }
    if (total > d_raw_cap_) {
      RN_CUDA(cudaStreamSynchronize(compute_));
      if (d_raw_) cudaFree(d_raw_);
      d_raw_ = nullptr;
      d_raw_cap_ = 0;
      RN_CUDA(cudaMalloc(reinterpret_cast<void**>(&d_raw_), total));
      d_raw_cap_ = total;
    }
    for (int i = 0; i < m; ++i)
      RN_CUDA(cudaMemcpyAsync(d_raw_ + descs[i].offset, imgs[off + i], static_cast<size_t>(H[off + i]) * W[off + i] * 3,
                              cudaMemcpyHostToDevice, compute_));
    RN_CUDA(cudaMemcpyAsync(d_descs_, descs.data(), m * sizeof(CropDesc), cudaMemcpyHostToDevice, compute_));
    RN_CUDA(CropResizeBatchU8(d_raw_, static_cast<const CropDesc*>(d_descs_), m, static_cast<uint8_t*>(d_in_[0]), S,
                              compute_));
    ++last_launches_;
    cur_ = &sets_[0];
    const HostOut ho = Out(0);
    cudaError_t e = ForwardDevice(d_in_[0], InputKind::kU8Bgr, m, ho.top1, ho.probs, ho.logits, compute_);
    if (e != cudaSuccess) return e;
    RN_CUDA(cudaStreamSynchronize(compute_));  // `descs` and the caller's images are pageable host memory
    Deliver(0, m, top1 ? top1 + off : nullptr, probs ? probs + static_cast<size_t>(off) * C : nullptr,
            logits ? logits + static_cast<size_t>(off) * C : nullptr);
  }
  return cudaSuccess;
}

cudaError_t Replica::InferYuv420(const uint8_t* y, const uint8_t* u, const uint8_t* v, size_t y_size, size_t u_size,
                                 size_t v_size, const YuvFrame& f, int64_t* top1, float* probs, float* logits,
                                 uint8_t* rgb_out) {
  {
    cudaError_t ew = WaitHost(~0ull);
    if (ew != cudaSuccess) return ew;
  }
  RN_CUDA(cudaSetDevice(device_));
  const int S = shape_.im_side;
  const size_t oy = 0, ou = (y_size + 255) & ~static_cast<size_t>(255), ov = ou + ((u_size + 255) & ~static_cast<size_t>(255));
  const size_t total = ov + v_size;
  if (total > d_raw_cap_) {
    RN_CUDA(cudaStreamSynchronize(compute_));
    if (d_raw_) cudaFree(d_raw_);
    d_raw_ = nullptr;
    d_raw_cap_ = 0;
    RN_CUDA(cudaMalloc(reinterpret_cast<void**>(&d_raw_), total));
    d_raw_cap_ = total;
  }
  RN_CUDA(cudaMemcpyAsync(d_raw_ + oy, y, y_size, cudaMemcpyHostToDevice, compute_));
  RN_CUDA(cudaMemcpyAsync(d_raw_ + ou, u, u_size, cudaMemcpyHostToDevice, compute_));
  RN_CUDA(cudaMemcpyAsync(d_raw_ + ov, v, v_size, cudaMemcpyHostToDevice, compute_));
  uint8_t* dst = static_cast<uint8_t*>(d_in_[0]);
  RN_CUDA(Yuv420CropU8(d_raw_ + oy, d_raw_ + ou, d_raw_ + ov, f, dst, S, compute_));
  last_launches_ = 1;
  if (rgb_out) RN_CUDA(cudaMemcpyAsync(rgb_out, dst, static_cast<size_t>(S) * S * 3, cudaMemcpyDeviceToHost, compute_));
  cur_ = &sets_[0];
  const HostOut ho = Out(0);
  cudaError_t e = ForwardDevice(dst, InputKind::kU8Rgb, 1, ho.top1, ho.probs, ho.logits, compute_);
  if (e != cudaSuccess) return e;
  RN_CUDA(cudaStreamSynchronize(compute_));
  Deliver(0, 1, top1, probs, logits);
  return cudaSuccess;
}

cudaError_t Replica::InferImage(const uint8_t* h_img, int H, int W, int64_t* top1, float* probs, float* logits) {
  cudaError_t e = Preprocess(h_img, H, W, nullptr);
  if (e != cudaSuccess) return e;
  last_launches_ = 1;
  cur_ = &sets_[0];
  const HostOut ho = Out(0);
  e = ForwardDevice(d_in_[0], InputKind::kU8Bgr, 1, ho.top1, ho.probs, ho.logits, compute_);
  if (e != cudaSuccess) return e;
  RN_CUDA(cudaStreamSynchronize(compute_));
  Deliver(0, 1, top1, probs, logits);
  return cudaSuccess;
}

cudaError_t Replica::DebugActivation(int layer, std::vector<float>* out, int dims[4]) {
  RN_CUDA(cudaSetDevice(device_));
  if (layer < 0 || layer >= kNumConvs || cur_->last_n <= 0) {
    err_ = "no activation recorded for that layer";
    return cudaErrorInvalidValue;
  }
  if (layer == 2 && block2_fused_last_ && precision_ != RN_PREC_FP32) {
    err_ = "conv2d_2's output stays on-chip in the fused residual-block kernel; create the handle with RN_FLAG_LAYERWISE";
    return cudaErrorInvalidValue;
  }
  const ConvShape& cs = shape_.conv[layer];
  dims[0] = cur_->last_n;
  dims[1] = dims[2] = cs.out_side;
  dims[3] = cs.cout;
  size_t elems = static_cast<size_t>(cur_->last_n) * cs.out_side * cs.out_side * cs.cout;
  out->resize(elems);
  RN_CUDA(cudaDeviceSynchronize());
  const float* src = nullptr;
  float* tmp = nullptr;
  if (layer >= first_f32_layer_) {
    src = cs.join_src >= 0 ? cur_->joined[layer] : cur_->pooled[layer];
  } else {
    // split tensors: conv0's eight channels are stored padded to sixteen (see Init)
    const int stored_ch = split_ ? std::max(cs.cout, 16) : cs.cout;
    const size_t stored_elems = elems / cs.cout * stored_ch;
    RN_CUDA(cudaMalloc(reinterpret_cast<void**>(&tmp), stored_elems * sizeof(float)));
    const void* h = cs.join_src >= 0 ? cur_->join_h[layer] : cur_->act_h[layer];
    const float sc = cs.join_src >= 0 ? 1.f : static_cast<float>(1.0 / act_scale_[layer]);
    cudaError_t e = ChunkedToF32(h, tmp, cur_->last_n, cs.out_side, stored_ch, half_kind_, sc, compute_, split_);
    if (e == cudaSuccess) e = cudaStreamSynchronize(compute_);
    if (e != cudaSuccess) {
      cudaFree(tmp);
      RN_CUDA(e);
    }
    if (stored_ch != cs.cout) {
      std::vector<float> wide(stored_elems);
      e = cudaMemcpy(wide.data(), tmp, stored_elems * sizeof(float), cudaMemcpyDeviceToHost);
      cudaFree(tmp);
      RN_CUDA(e);
      for (size_t px = 0; px < elems / cs.cout; ++px)
        for (int c = 0; c < cs.cout; ++c) (*out)[px * cs.cout + c] = wide[px * stored_ch + c];
      return cudaSuccess;
    }
    src = tmp;
  }
  cudaError_t e = cudaMemcpy(out->data(), src, elems * sizeof(float), cudaMemcpyDeviceToHost);
  if (tmp) cudaFree(tmp);
  RN_CUDA(e);
  if (layer < first_f32_layer_ && !join_gain_[layer].empty())  // stored with a per-channel gain (LoadFolded)
    for (size_t k = 0; k < elems; ++k) (*out)[k] = static_cast<float>((*out)[k] / join_gain_[layer][k % cs.cout]);
  return cudaSuccess;
}

}  // namespace rn
