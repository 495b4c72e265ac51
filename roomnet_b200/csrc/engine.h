// One inference replica = one GPU: folded weights, resident activation buffers,
// pinned staging, two streams.  Replaces what tf.Session held for the reference
// (network.py:87-91) on a single device.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <string>
#include <vector>

#include "fold.h"
#include "jpeg_host.h"
#include "kernels.h"

namespace rn {

// kArgb8888: one 32-bit 0xAARRGGBB int per pixel (android.graphics.Bitmap.getPixels), i.e. bytes B,G,R,A in memory
enum class InputKind : int { kU8Bgr = 0, kU8Rgb = 1, kF32Rgb = 2, kArgb8888 = 3 };
size_t InputBytesPerPixel(InputKind k);

class Replica {
 public:
  Replica(int device, const NetShape& shape, int precision, int max_batch, int flags = 0);
  ~Replica();
  Replica(const Replica&) = delete;
  Replica& operator=(const Replica&) = delete;

  cudaError_t Init();
  cudaError_t Upload(const FoldedNet& f);
  bool loaded() const { return loaded_; }
  int device() const { return device_; }
  int max_batch() const { return max_batch_; }

  // n <= max_batch images already resident on this device; enqueues on `st` using activation set `cur_`.
  cudaError_t ForwardDevice(const void* d_in, InputKind kind, int n, long long* d_top1, float* d_probs,
                            float* d_logits, cudaStream_t st);
  // Arbitrary n from host memory (pinned or pageable), double-buffered micro-batches; synchronous.
  cudaError_t InferHost(const void* h_in, InputKind kind, int n, int64_t* top1, float* probs, float* logits);
  // n BGR uint8 photos of arbitrary sizes (imgs[i] = H[i] x W[i] x 3): centre crop + cv2-identical resize of a whole
  // micro-batch in one launch, straight into the network input - no host round trip between the two.
  // One YUV_420_888 camera frame: colour conversion + frame-to-crop transform + forward pass on the device.
  cudaError_t InferYuv420(const uint8_t* y, const uint8_t* u, const uint8_t* v, size_t y_size, size_t u_size,
                          size_t v_size, const YuvFrame& f, int64_t* top1, float* probs, float* logits,
                          uint8_t* rgb_out);
  cudaError_t InferImages(const uint8_t* const* imgs, const int* H, const int* W, int n, int64_t* top1, float* probs,
                          float* logits);
  // n encoded baseline-JPEG files (cv2.imread + infer_optimized of infer.py:81-82): entropy decoding on `threads` host
  // threads, everything else on the device.  status[i] = JpegStatus; files that are not kJpegOk leave their outputs
  // untouched (the caller decodes those on the host).
  cudaError_t InferJpegs(const uint8_t* const* files, const size_t* sizes, int n, int threads, int64_t* top1, float* probs,
                         float* logits, int32_t* status);
  // Decode only: the BGR image as cv2.imread returns it (EXIF orientation applied); out == nullptr queries the size.
  cudaError_t DecodeJpeg(const uint8_t* file, size_t size, uint8_t* out, size_t capacity, int* height, int* width,
                         int32_t* status);
  // Asynchronous form: SubmitHost enqueues the copies and kernels of one call and returns (it blocks only when every
  // staging slot is still in flight); the results are delivered to the caller's buffers by a later SubmitHost that
  // needs the slot, or by WaitHost(ticket), which returns once every call up to `ticket` has been delivered.
  cudaError_t SubmitHost(const void* h_in, InputKind kind, int n, int64_t* top1, float* probs, float* logits,
                         uint64_t ticket);
  cudaError_t WaitHost(uint64_t ticket);
  // Arbitrary n, device-resident input/outputs, micro-batched on `st` (nullptr = own stream); no sync.
  cudaError_t InferDevice(const void* d_in, InputKind kind, int n, long long* d_top1, float* d_probs,
                          float* d_logits, cudaStream_t st);

  // center_crop + cv2.resize(im, (S, S)) of one HxWx3 uint8 image on the device (reference network.py:149-152).
  // Writes S*S*3 bytes to h_out (host, may be null) and leaves the result in the replica's input slot 0.
  cudaError_t Preprocess(const uint8_t* h_img, int H, int W, uint8_t* h_out);
  // infer_optimized for one arbitrary-size BGR image entirely on the device (Preprocess + forward).
  cudaError_t InferImage(const uint8_t* h_img, int H, int W, int64_t* top1, float* probs, float* logits);
  int last_launches() const { return last_launches_; }
  // files whose Huffman decoding ran on the device / on host threads so far
  void jpeg_counters(long long* device_files, long long* host_files) const {
    *device_files = jpeg_device_huffman_files_;
    *host_files = jpeg_host_huffman_files_;
  }
  // Per-kernel device timing (CUDA events on the launching stream, recorded between launches).
  void set_profiling(bool on) { profiling_ = on; }
  struct KernelTime {
    std::string name;
    double ms = 0;
    int launches = 0;
  };
  cudaError_t ProfileResults(std::vector<KernelTime>* out);
  // Debug: copies the output of conv layer `layer` (pooled, after the residual join where there is
  // one) of the last micro-batch to host as NHWC fp32.
  cudaError_t DebugActivation(int layer, std::vector<float>* out, int dims[4]);
  const std::string& error() const { return err_; }

 private:
  cudaError_t ForwardF32(const void* d_in, InputKind kind, int n, cudaStream_t st);
  cudaError_t ForwardTc(const void* d_in, InputKind kind, int n, long long* d_top1, float* d_probs, float* d_logits,
                        cudaStream_t st);
  cudaError_t DenseTail(int n, long long* d_top1, float* d_probs, float* d_logits, cudaStream_t st);
  cudaError_t TailF32(int first_layer, int n, cudaStream_t st);
  cudaError_t Alloc(void** p, size_t bytes);
  cudaError_t UploadF32(const std::vector<double>& v, float** dptr);

  int device_, precision_, max_batch_;
  NetShape shape_;
  bool loaded_ = false;
  std::string err_;
  int last_launches_ = 0;
  bool profiling_ = false;
  std::vector<cudaEvent_t> prof_events_;
  std::vector<std::string> prof_names_;  // prof_names_[i] = kernel between event i and i+1 ("" = gap)
  size_t prof_used_ = 0;
  void Mark(const char* name, cudaStream_t st);

  cudaStream_t compute_ = nullptr, copy_ = nullptr;
  // Host entry points: a ring of kSlots staging slots (device input buffer, device/host result buffers, events) feeds
  // the two activation sets (micro-batch j: slot j % kSlots, set j & 1).  With more slots than sets the host->device
  // copies run up to kSlots - 1 micro-batches ahead of the kernels, so a stream of calls is bound by
  // max(PCIe time, kernel time) instead of their partial sum.
  static constexpr int kSlots = 8;
  cudaEvent_t ev_h2d_[kSlots] = {}, ev_done_[kSlots] = {};
  struct PendingOut {  // a micro-batch in flight in staging slot s: where its results go once ev_done_[s] has fired
    int m = 0;
    bool active = false;
    int64_t* top1 = nullptr;
    float* probs = nullptr;
    float* logits = nullptr;
    uint64_t ticket = 0;
  } pend_[kSlots];
  unsigned slot_seq_ = 0;  // micro-batches submitted so far
  cudaError_t DrainSlot(int slot);
  void AbortPending();
  std::vector<void*> allocs_;

  // weights (fp32, HWIO folded)
  float* w0_[3] = {nullptr, nullptr, nullptr};  // conv0 per InputKind
  float* b0_[3] = {nullptr, nullptr, nullptr};
  float* cw_[kNumConvs] = {};
  float* cb_[kNumConvs] = {};
  float* ja_[kNumConvs] = {};
  float* jb_[kNumConvs] = {};
  float* jc_[kNumConvs] = {};
  DenseParams dense_{};
  // 16-bit packed weights
  TcConvLayer tc_[kNumConvs] = {};
  // 16-bit path bookkeeping: stored activation = true value * act_scale_[i]  (k*k/6 after a pooled
  // tensor-core layer, 1 for conv0 and for residual-join outputs)
  double act_scale_[kNumConvs] = {};
  float* tc_bias_[kNumConvs] = {};
  void* tc0_w_[2] = {nullptr, nullptr};  // conv0 packed weights for uint8 BGR / RGB byte order
  float* tc_abc_[kNumConvs] = {};  // [3][cout] A/B/C for the join fused into the conv epilogue
  std::vector<double> join_gain_[kNumConvs];  // per-channel gain a join output is stored with (empty = none)
  HalfKind half_kind_ = HalfKind::kF16;

  // Activations.  Two independent sets, each with its own stream: consecutive micro-batches alternate between
  // them so that the memory-bound kernels of one (residual joins, image expansion, tail) overlap the
  // tensor-core kernels of the other on the same SMs.
  struct ActSet {
    float* conv_scratch = nullptr;
    float* pooled[kNumConvs] = {};
    float* joined[kNumConvs] = {};
    void* act_h[kNumConvs] = {};
    void* join_h[kNumConvs] = {};
    void* in_h = nullptr;  // uint8 image expanded to pixel-pair chunks (PrepU8)
    float* d_pre = nullptr;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev_done = nullptr;
    int last_n = 0;
  };
  ActSet sets_[2];
  ActSet* cur_ = &sets_[0];
  cudaEvent_t ev_fork_ = nullptr;
  int first_f32_layer_ = 0;  // layers >= this run on the fp32 kernels
  bool split_ = false;       // RN_PREC_FP32_TC: three-product split-fp16 tensor-core layers, fp32 epilogues
  bool layerwise_ = false;   // RN_FLAG_LAYERWISE: no fused residual-block kernel
  bool block2_fused_last_ = false;  // the last forward pass ran residual block 2 as one kernel

  // JPEG front end: pinned coefficient staging, device coefficient / sample-plane arenas, descriptors; grown on demand
  struct JpegBatch;
  cudaError_t InferJpegsHostHuffman(const uint8_t* const* files, const size_t* sizes, int n, int threads, int64_t* top1,
                                    float* probs, float* logits, int32_t* status);
  cudaError_t JpegHuffUpload(const uint8_t* const* files, const size_t* sizes, JpegBatch* b, int threads,
                             const uint8_t* h_prepared, int slot);
  cudaError_t JpegHuffRun(JpegBatch* b, int slot, std::vector<int>* err);
  struct HuffStage {  // a batch between its upload (copy stream) and its kernels (compute stream)
    HuffBatch hb{};
    std::vector<int> file_of;
    size_t o_err = 0, coef_bytes = 0;
    int nf = 0;
    bool any = false;
  } huff_stage_[2];
  cudaEvent_t ev_huff_up_[2] = {nullptr, nullptr};
  void JpegHuffmanErrors(std::vector<int>* err) const;
  uint8_t* d_huff_[2] = {nullptr, nullptr};  // device arenas of the Huffman stage (streams, per-subsequence state,
  size_t d_huff_cap_[2] = {0, 0};             // descriptors): one is uploaded to while the kernels read the other
  int* h_huff_flags_ = nullptr;  // pinned: [0] convergence flag, [16 ..] per-file error flags
  std::vector<int> huff_file_of_;
  int jpeg_huffman_rounds_ = 0;
  long long jpeg_device_huffman_files_ = 0, jpeg_host_huffman_files_ = 0;
  int flags_ = 0;
  cudaError_t GrowJpegBuffers(size_t coef_bytes, size_t sample_bytes, size_t raw_bytes, int n_images, int n_host);
  void JpegDecodeHost(const uint8_t* const* files, const size_t* sizes, JpegBatch* b, int16_t* h_coef, int threads);
  cudaError_t JpegToRaw(const JpegBatch& b, const int16_t* h_coef, std::vector<CropDesc>* crops, std::vector<char>* ok,
                        int32_t* status);
  static constexpr int kJpegRing = 4;
  int16_t* h_coef_[kJpegRing] = {};  // pinned ring: host threads fill buffers ahead of the one the device reads
  int16_t* d_coef_ = nullptr;
  uint8_t* d_samples_ = nullptr;
  void* d_jmeta_ = nullptr;
  size_t h_coef_cap_[kJpegRing] = {}, d_coef_cap_ = 0, d_samples_cap_ = 0, d_jmeta_cap_ = 0;
  // preprocessing scratch (raw image + tap tables), grown on demand
  uint8_t* d_raw_ = nullptr;
  size_t d_raw_cap_ = 0;
  int* d_taps_ = nullptr;
  void* d_descs_ = nullptr;  // CropDesc[max_batch] of the batched preprocess
  // staging
  void* d_in_[kSlots] = {};
  void* h_in_[2] = {nullptr, nullptr};  // pinned bounce buffers for pageable caller memory (micro-batch j: j & 1)
  char* h_out_[kSlots] = {};  // pinned + device-mapped result buffers, written by the tail kernel itself
  struct HostOut {
    long long* top1;
    float* probs;
    float* logits;
  };
  HostOut Out(int slot) const;
  void Deliver(int slot, int m, int64_t* top1, float* probs, float* logits) const;
};

}  // namespace rn
