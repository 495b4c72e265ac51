// C ABI façade + replica scheduler (see include/roomnet.h for the reference
// interface each entry point replaces).
#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <cstring>
#include <functional>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "engine.h"
#include "fold.h"
#include "roomnet.h"
#include "tf_bundle.h"

using rn::InputKind;

namespace {
// One persistent host worker per replica (GPU): a call hands every worker its contiguous shard and waits for all
// of them; no thread is created on the inference path (one host worker thread + streams per GPU).
class Worker {
 public:
  Worker() : th_([this] { Loop(); }) {}
  ~Worker() {
    {
      std::lock_guard<std::mutex> l(mu_);
      stop_ = true;
    }
    cv_.notify_all();
    th_.join();
  }
  void Submit(std::function<void()> fn) {
    {
      std::lock_guard<std::mutex> l(mu_);
      task_ = std::move(fn);
      busy_ = true;
    }
    cv_.notify_all();
  }
  void Wait() {
    std::unique_lock<std::mutex> l(mu_);
    done_cv_.wait(l, [this] { return !busy_; });
  }

 private:
  void Loop() {
    std::unique_lock<std::mutex> l(mu_);
    for (;;) {
      cv_.wait(l, [this] { return stop_ || task_; });
      if (stop_) return;
      std::function<void()> fn = std::move(task_);
      task_ = nullptr;
      l.unlock();
      fn();
      l.lock();
      busy_ = false;
      done_cv_.notify_all();
    }
  }
  std::mutex mu_;
  std::condition_variable cv_, done_cv_;
  std::function<void()> task_;
  bool busy_ = false, stop_ = false;
  std::thread th_;
};
}  // namespace

struct rn_handle {
  rn_config cfg{};
  rn::NetShape shape{};
  std::vector<std::unique_ptr<rn::Replica>> replicas;
  std::vector<std::unique_ptr<Worker>> workers;  // one per replica when there are several (declared after: joins first)
  rn::FoldedNet folded;
  bool loaded = false;
  bool has_dense0 = false;
  rn::Tensor dense0;
  std::string err;
  std::mutex mu;
  int last_launches = 0;
  std::vector<double> lat_ms;
  int64_t calls = 0, images = 0;
  uint64_t next_ticket = 1;                 // rn_submit_*: tickets are handed out in submission order
  std::vector<cudaError_t> async_status;    // per replica: first error of a submitted call (reported by rn_wait)
};

namespace {

thread_local std::string g_create_error;

int Fail(rn_handle* h, int code, const std::string& msg) {
  if (h) h->err = msg;
  return code;
}

// No exception crosses the C boundary (std::bad_alloc from the vectors, std::system_error from the worker threads):
// every entry point that allocates runs its body through this.
template <typename Fn>
int Guarded(rn_handle* h, Fn&& fn) {
  try {
    return fn();
  } catch (const std::exception& e) {
    return Fail(h, RN_ERR_INTERNAL, std::string("internal error: ") + e.what());
  } catch (...) {
    return Fail(h, RN_ERR_INTERNAL, "internal error");
  }
}

int LoadCommon(rn_handle* h, const rn::TensorMap& vars) {
  std::string err;
  if (!rn::FoldNetwork(vars, h->shape, h->has_dense0 ? &h->dense0 : nullptr, &h->folded, &err))
    return Fail(h, RN_ERR_FORMAT, err);
  for (auto& r : h->replicas) {
    if (r->Upload(h->folded) != cudaSuccess) return Fail(h, RN_ERR_CUDA, r->error());
  }
  h->loaded = true;
  return RN_OK;
}

int InferImpl(rn_handle* h, const void* in, InputKind kind, int32_t n, int64_t* top1, float* probs, float* logits) {
  if (!h) return RN_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lock(h->mu);
  if (!in || n < 0) return Fail(h, RN_ERR_INVALID_ARG, "null input or negative batch size");
  if (h->replicas.empty()) return Fail(h, RN_ERR_CUDA, "host-only handle (n_devices = 0): there is no CPU inference path");
  if (!h->loaded) return Fail(h, RN_ERR_NOT_LOADED, "weights have not been loaded (rn_load_tf_checkpoint)");
  if (n == 0) return RN_OK;
  auto t0 = std::chrono::steady_clock::now();
  const int g = static_cast<int>(h->replicas.size());
  const int C = h->shape.num_classes;
  const size_t per = static_cast<size_t>(h->shape.im_side) * h->shape.im_side * rn::InputBytesPerPixel(kind);
  // contiguous, as-even-as-possible split of [0, n) over the replicas
  std::vector<int> begin(g + 1, 0);
  for (int r = 0; r < g; ++r) begin[r + 1] = begin[r] + n / g + (r < n % g ? 1 : 0);
  std::vector<cudaError_t> status(g, cudaSuccess);
  auto run = [&](int r) {
    int b = begin[r], m = begin[r + 1] - begin[r];
    if (m == 0) return;
    status[r] = h->replicas[r]->InferHost(static_cast<const char*>(in) + per * b, kind, m, top1 ? top1 + b : nullptr,
                                          probs ? probs + static_cast<size_t>(b) * C : nullptr,
                                          logits ? logits + static_cast<size_t>(b) * C : nullptr);
  };
  if (g == 1) {
    run(0);
  } else {
    for (int r = 0; r < g; ++r) {
      h->workers[r]->Wait();  // a shard submitted through rn_submit_* may still be enqueuing
      h->workers[r]->Submit([&run, r] { run(r); });
    }
    for (int r = 0; r < g; ++r) h->workers[r]->Wait();
  }
  h->last_launches = 0;
  for (int r = 0; r < g; ++r) {
    if (status[r] != cudaSuccess) return Fail(h, RN_ERR_CUDA, h->replicas[r]->error());
    h->last_launches += h->replicas[r]->last_launches();
  }
  double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  if (h->lat_ms.size() < (1u << 20)) h->lat_ms.push_back(ms);
  h->calls += 1;
  h->images += n;
  return RN_OK;
}

// rn_submit_*: the shards of one call are enqueued by the replicas' workers (one replica: by the caller); nothing waits
// for the GPU here unless every staging slot of a replica (Replica::kSlots micro-batches) is still in flight.
int SubmitImpl(rn_handle* h, const void* in, InputKind kind, int32_t n, int64_t* top1, float* probs, float* logits,
               uint64_t* ticket) {
  if (!h) return RN_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lock(h->mu);
  if (!in || n < 0 || !ticket) return Fail(h, RN_ERR_INVALID_ARG, "null input / ticket or negative batch size");
  if (h->replicas.empty()) return Fail(h, RN_ERR_CUDA, "host-only handle (n_devices = 0): there is no CPU inference path");
  if (!h->loaded) return Fail(h, RN_ERR_NOT_LOADED, "weights have not been loaded (rn_load_tf_checkpoint)");
  const uint64_t tk = h->next_ticket++;
  *ticket = tk;
  if (n == 0) return RN_OK;
  const int g = static_cast<int>(h->replicas.size());
  const int C = h->shape.num_classes;
  const size_t per = static_cast<size_t>(h->shape.im_side) * h->shape.im_side * rn::InputBytesPerPixel(kind);
  h->async_status.resize(g, cudaSuccess);
  int b = 0;
  for (int r = 0; r < g; ++r) {
    const int m = n / g + (r < n % g ? 1 : 0);
    if (m == 0) continue;
    rn::Replica* rep = h->replicas[r].get();
    cudaError_t* status = &h->async_status[r];
    auto task = [=] {
      cudaError_t e = rep->SubmitHost(static_cast<const char*>(in) + per * b, kind, m, top1 ? top1 + b : nullptr,
                                      probs ? probs + static_cast<size_t>(b) * C : nullptr,
                                      logits ? logits + static_cast<size_t>(b) * C : nullptr, tk);
      if (e != cudaSuccess && *status == cudaSuccess) *status = e;
    };
    if (g == 1) {
      task();
    } else {
      h->workers[r]->Wait();  // the previous call's shard has been enqueued
      h->workers[r]->Submit(task);
    }
    b += m;
  }
  h->calls += 1;
  h->images += n;
  if (g == 1 && h->async_status[0] != cudaSuccess) {
    h->async_status[0] = cudaSuccess;
    return Fail(h, RN_ERR_CUDA, h->replicas[0]->error());
  }
  return RN_OK;
}

int WaitImpl(rn_handle* h, uint64_t ticket) {
  if (!h) return RN_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lock(h->mu);
  if (ticket == 0) ticket = ~0ull;
  const int g = static_cast<int>(h->replicas.size());
  int rc = RN_OK;
  for (int r = 0; r < g; ++r) {
    if (g > 1) h->workers[r]->Wait();
    cudaError_t e = h->replicas[r]->WaitHost(ticket);
    if (r < static_cast<int>(h->async_status.size()) && h->async_status[r] != cudaSuccess) {
      e = h->async_status[r];
      h->async_status[r] = cudaSuccess;
    }
    if (e != cudaSuccess && rc == RN_OK) rc = Fail(h, RN_ERR_CUDA, h->replicas[r]->error());
  }
  return rc;
}

int Infer(rn_handle* h, const void* in, InputKind kind, int32_t n, int64_t* top1, float* probs, float* logits) {
  return Guarded(h, [&] { return InferImpl(h, in, kind, n, top1, probs, logits); });
}

}  // namespace

extern "C" {

int rn_submit_u8_bgr(rn_handle* h, const uint8_t* nhwc, int32_t n, int64_t* top1, float* probs, float* logits,
                     uint64_t* ticket) {
  return Guarded(h, [&] { return SubmitImpl(h, nhwc, InputKind::kU8Bgr, n, top1, probs, logits, ticket); });
}
int rn_wait(rn_handle* h, uint64_t ticket) {
  return Guarded(h, [&] { return WaitImpl(h, ticket); });
}

const char* rn_version(void) { return "roomnet_b200 0.1 (sm_100a)"; }

const char* rn_last_error(const rn_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

static int CreateImpl(const rn_config* cfg, rn_handle** out) {
  g_create_error.clear();
  if (!cfg || !out) {
    g_create_error = "null config or output pointer";
    return RN_ERR_INVALID_ARG;
  }
  *out = nullptr;
  if (cfg->abi_version != RN_ABI_VERSION) {
    g_create_error = "rn_config.abi_version mismatch";
    return RN_ERR_INVALID_ARG;
  }
  if (cfg->precision < RN_PREC_FP32 || cfg->precision > RN_PREC_BF16X3) {
    g_create_error = "unknown precision";
    return RN_ERR_INVALID_ARG;
  }
  // n_devices == 0 makes a host-only handle: checkpoint parsing and BN folding work
  // (rn_load_*, rn_get_folded), every inference entry point fails with RN_ERR_CUDA.
  if (cfg->n_devices < 0 || cfg->n_devices > RN_MAX_DEVICES) {
    g_create_error = "n_devices out of range";
    return RN_ERR_INVALID_ARG;
  }
  if (cfg->num_classes < 1 || cfg->num_classes > 32) {
    g_create_error = "num_classes out of range (1..32)";
    return RN_ERR_INVALID_ARG;
  }
  auto h = std::make_unique<rn_handle>();
  h->cfg = *cfg;
  std::string err;
  if (!rn::MakeNetShape(cfg->im_side, cfg->num_classes, &h->shape, &err)) {
    g_create_error = err;
    return RN_ERR_INVALID_ARG;
  }
  if (cfg->max_batch > 4096) {
    g_create_error = "max_batch out of range (<= 4096 images resident per replica; larger calls are micro-batched anyway)";
    return RN_ERR_INVALID_ARG;
  }
  if (cfg->flags & ~(RN_FLAG_LAYERWISE | RN_FLAG_JPEG_HOST_HUFFMAN)) {
    g_create_error = "unknown bits in rn_config.flags";
    return RN_ERR_INVALID_ARG;
  }
  int mb = cfg->max_batch;
  if (mb <= 0) {
    // resident micro-batch: bounded by activation memory, which grows with im_side^2
    double scale = (224.0 * 224.0) / (static_cast<double>(cfg->im_side) * cfg->im_side);
    mb = std::max(1, std::min(256, static_cast<int>(256 * scale)));
  }
  h->cfg.max_batch = mb;
  for (int i = 0; i < cfg->n_devices; ++i) {
    auto r = std::make_unique<rn::Replica>(cfg->devices[i], h->shape, cfg->precision, mb, cfg->flags);
    if (r->Init() != cudaSuccess) {
      g_create_error = r->error();
      return RN_ERR_CUDA;
    }
    h->replicas.push_back(std::move(r));
  }
  if (cfg->n_devices > 1)
    for (int i = 0; i < cfg->n_devices; ++i) h->workers.push_back(std::make_unique<Worker>());
  *out = h.release();
  return RN_OK;
}

int rn_destroy(rn_handle* h) {
  delete h;
  return RN_OK;
}

static int LoadCheckpointImpl(rn_handle* h, const char* prefix) {
  if (!h) return RN_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lock(h->mu);
  if (!prefix) return Fail(h, RN_ERR_INVALID_ARG, "null checkpoint prefix");
  rn::TensorMap vars;
  std::string err;
  rn::BundleError be = rn::ReadBundle(prefix, &vars, &err);
  if (be == rn::BundleError::kIo) return Fail(h, RN_ERR_IO, err);
  if (be != rn::BundleError::kOk) return Fail(h, RN_ERR_FORMAT, err);
  return LoadCommon(h, vars);
}

static int LoadTensorsImpl(rn_handle* h, int32_t n, const char* const* names, const float* const* data,
                           const int64_t* const* shapes, const int32_t* ranks) {
  if (!h) return RN_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lock(h->mu);
  if (n < 0 || !names || !data || !shapes || !ranks) return Fail(h, RN_ERR_INVALID_ARG, "null argument");
  rn::TensorMap vars;
  for (int i = 0; i < n; ++i) {
    if (!names[i] || !data[i] || ranks[i] < 0 || ranks[i] > 8 || (ranks[i] && !shapes[i]))
      return Fail(h, RN_ERR_INVALID_ARG, "bad tensor entry " + std::to_string(i));
    rn::Tensor t;
    t.shape.assign(shapes[i], shapes[i] + ranks[i]);
    for (auto d : t.shape)
      if (d < 0) return Fail(h, RN_ERR_INVALID_ARG, "negative dimension");
    t.data.assign(data[i], data[i] + t.numel());
    vars[names[i]] = std::move(t);
  }
  return LoadCommon(h, vars);
}

int rn_create(const rn_config* cfg, rn_handle** out) {
  try {
    return CreateImpl(cfg, out);
  } catch (const std::exception& e) {
    g_create_error = std::string("internal error: ") + e.what();
  } catch (...) {
    g_create_error = "internal error";
  }
  if (out) *out = nullptr;
  return RN_ERR_INTERNAL;
}
int rn_load_tf_checkpoint(rn_handle* h, const char* prefix) {
  return Guarded(h, [&] { return LoadCheckpointImpl(h, prefix); });
}
int rn_load_tensors(rn_handle* h, int32_t n, const char* const* names, const float* const* data,
                    const int64_t* const* shapes, const int32_t* ranks) {
  return Guarded(h, [&] { return LoadTensorsImpl(h, n, names, data, shapes, ranks); });
}

int rn_set_dense0(rn_handle* h, const float* kernel, int32_t flat_len) {
  if (!h) return RN_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lock(h->mu);
  if (!kernel) return Fail(h, RN_ERR_INVALID_ARG, "null kernel");
  if (flat_len != h->shape.flat_len)
    return Fail(h, RN_ERR_INVALID_ARG,
                "dense/kernel override has " + std::to_string(flat_len) + " rows but im_side " +
                    std::to_string(h->shape.im_side) + " flattens to " + std::to_string(h->shape.flat_len));
  h->dense0.shape = {flat_len, h->shape.dense_out[0]};
  h->dense0.data.assign(kernel, kernel + static_cast<size_t>(flat_len) * h->shape.dense_out[0]);
  h->has_dense0 = true;
  return RN_OK;
}

int rn_infer_u8_bgr(rn_handle* h, const uint8_t* nhwc, int32_t n, int64_t* top1, float* probs, float* logits) {
  return Infer(h, nhwc, InputKind::kU8Bgr, n, top1, probs, logits);
}
int rn_infer_u8_rgb(rn_handle* h, const uint8_t* nhwc, int32_t n, int64_t* top1, float* probs, float* logits) {
  return Infer(h, nhwc, InputKind::kU8Rgb, n, top1, probs, logits);
}
int rn_infer_f32_rgb(rn_handle* h, const float* nhwc, int32_t n, int64_t* top1, float* probs, float* logits) {
  return Infer(h, nhwc, InputKind::kF32Rgb, n, top1, probs, logits);
}
int rn_infer_argb8888(rn_handle* h, const int32_t* pixels, int32_t n, int64_t* top1, float* probs, float* logits) {
  return Infer(h, pixels, InputKind::kArgb8888, n, top1, probs, logits);
}

int rn_infer_u8_bgr_device(rn_handle* h, const void* d_nhwc, int32_t n, void* d_top1, void* d_probs, void* d_logits,
                           void* cuda_stream) {
  if (!h) return RN_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lock(h->mu);
  if (!d_nhwc || n < 0) return Fail(h, RN_ERR_INVALID_ARG, "null input or negative batch size");
  if (h->replicas.empty()) return Fail(h, RN_ERR_CUDA, "host-only handle (n_devices = 0): there is no CPU inference path");
  if (!h->loaded) return Fail(h, RN_ERR_NOT_LOADED, "weights have not been loaded (rn_load_tf_checkpoint)");
  if (n == 0) return RN_OK;
  rn::Replica* r = h->replicas[0].get();
  cudaError_t e = r->InferDevice(d_nhwc, InputKind::kU8Bgr, n, static_cast<long long*>(d_top1),
                                 static_cast<float*>(d_probs), static_cast<float*>(d_logits),
                                 static_cast<cudaStream_t>(cuda_stream));
  if (e != cudaSuccess) return Fail(h, RN_ERR_CUDA, r->error());
  h->last_launches = r->last_launches();
  return RN_OK;
}

int rn_preprocess_u8(rn_handle* h, const uint8_t* img, int32_t hgt, int32_t wid, uint8_t* out) {
  if (!h) return RN_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lock(h->mu);
  if (!img || !out || hgt < 2 || wid < 2) return Fail(h, RN_ERR_INVALID_ARG, "null image or degenerate size");
  if (h->replicas.empty()) return Fail(h, RN_ERR_CUDA, "host-only handle (n_devices = 0): there is no CPU path");
  if (h->replicas[0]->Preprocess(img, hgt, wid, out) != cudaSuccess) return Fail(h, RN_ERR_CUDA, h->replicas[0]->error());
  return RN_OK;
}

int rn_infer_image_u8_bgr(rn_handle* h, const uint8_t* img, int32_t hgt, int32_t wid, int64_t* top1, float* probs,
                          float* logits) {
  if (!h) return RN_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lock(h->mu);
  if (!img || hgt < 2 || wid < 2) return Fail(h, RN_ERR_INVALID_ARG, "null image or degenerate size");
  if (h->replicas.empty()) return Fail(h, RN_ERR_CUDA, "host-only handle (n_devices = 0): there is no CPU inference path");
  if (!h->loaded) return Fail(h, RN_ERR_NOT_LOADED, "weights have not been loaded (rn_load_tf_checkpoint)");
  auto t0 = std::chrono::steady_clock::now();
  rn::Replica* r = h->replicas[0].get();
  if (r->InferImage(img, hgt, wid, top1, probs, logits) != cudaSuccess) return Fail(h, RN_ERR_CUDA, r->error());
  h->last_launches = r->last_launches();
  double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  if (h->lat_ms.size() < (1u << 20)) h->lat_ms.push_back(ms);
  h->calls += 1;
  h->images += 1;
  return RN_OK;
}

int rn_infer_images_u8_bgr(rn_handle* h, const uint8_t* const* imgs, const int32_t* heights, const int32_t* widths,
                           int32_t n, int64_t* top1, float* probs, float* logits) {
  return Guarded(h, [&]() -> int {
    if (!h) return RN_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> lock(h->mu);
    if (!imgs || !heights || !widths || n < 0) return Fail(h, RN_ERR_INVALID_ARG, "null argument or negative count");
    for (int i = 0; i < n; ++i)
      if (!imgs[i] || heights[i] < 2 || widths[i] < 2)
        return Fail(h, RN_ERR_INVALID_ARG, "image " + std::to_string(i) + ": null pointer or degenerate size");
    if (h->replicas.empty()) return Fail(h, RN_ERR_CUDA, "host-only handle (n_devices = 0): there is no CPU inference path");
    if (!h->loaded) return Fail(h, RN_ERR_NOT_LOADED, "weights have not been loaded (rn_load_tf_checkpoint)");
    if (n == 0) return RN_OK;
    auto t0 = std::chrono::steady_clock::now();
    const int g = static_cast<int>(h->replicas.size());
    const int C = h->shape.num_classes;
    std::vector<int> begin(g + 1, 0);
    for (int r = 0; r < g; ++r) begin[r + 1] = begin[r] + n / g + (r < n % g ? 1 : 0);
    std::vector<cudaError_t> status(g, cudaSuccess);
    auto run = [&](int r) {
      const int b = begin[r], m = begin[r + 1] - begin[r];
      if (m == 0) return;
      status[r] = h->replicas[r]->InferImages(imgs + b, heights + b, widths + b, m, top1 ? top1 + b : nullptr,
                                              probs ? probs + static_cast<size_t>(b) * C : nullptr,
                                              logits ? logits + static_cast<size_t>(b) * C : nullptr);
    };
    if (g == 1) {
      run(0);
    } else {
      for (int r = 0; r < g; ++r) {
        h->workers[r]->Wait();
        h->workers[r]->Submit([&run, r] { run(r); });
      }
      for (int r = 0; r < g; ++r) h->workers[r]->Wait();
    }
    h->last_launches = 0;
    for (int r = 0; r < g; ++r) {
      if (status[r] != cudaSuccess) return Fail(h, RN_ERR_CUDA, h->replicas[r]->error());
      h->last_launches += h->replicas[r]->last_launches();
    }
    double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    if (h->lat_ms.size() < (1u << 20)) h->lat_ms.push_back(ms);
    h->calls += 1;
    h->images += n;
    return RN_OK;
  });
}

int rn_infer_jpeg(rn_handle* h, const uint8_t* const* files, const uint64_t* sizes, int32_t n, int32_t threads,
                  int64_t* top1, float* probs, float* logits, int32_t* status) {
  return Guarded(h, [&]() -> int {
    if (!h) return RN_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> lock(h->mu);
    if (!files || !sizes || !status || n < 0) return Fail(h, RN_ERR_INVALID_ARG, "null argument or negative count");
    for (int i = 0; i < n; ++i)
      if (!files[i]) return Fail(h, RN_ERR_INVALID_ARG, "file " + std::to_string(i) + ": null pointer");
    if (h->replicas.empty()) return Fail(h, RN_ERR_CUDA, "host-only handle (n_devices = 0): there is no CPU inference path");
    if (!h->loaded) return Fail(h, RN_ERR_NOT_LOADED, "weights have not been loaded (rn_load_tf_checkpoint)");
    if (n == 0) return RN_OK;
    auto t0 = std::chrono::steady_clock::now();
    const int g = static_cast<int>(h->replicas.size());
    const int C = h->shape.num_classes;
    int hw = static_cast<int>(std::thread::hardware_concurrency());
    if (hw < 1) hw = 1;
    const int nt = std::max(1, (threads > 0 ? threads : std::min(hw, 16)) / g);
    std::vector<size_t> sz(sizes, sizes + n);
    std::vector<int> begin(g + 1, 0);
    for (int r = 0; r < g; ++r) begin[r + 1] = begin[r] + n / g + (r < n % g ? 1 : 0);
    std::vector<cudaError_t> st(g, cudaSuccess);
    auto run = [&](int r) {
      const int b = begin[r], m = begin[r + 1] - begin[r];
      if (m == 0) return;
      st[r] = h->replicas[r]->InferJpegs(files + b, sz.data() + b, m, nt, top1 ? top1 + b : nullptr,
                                         probs ? probs + static_cast<size_t>(b) * C : nullptr,
                                         logits ? logits + static_cast<size_t>(b) * C : nullptr, status + b);
    };
    if (g == 1) {
      run(0);
    } else {
      for (int r = 0; r < g; ++r) {
        h->workers[r]->Wait();
        h->workers[r]->Submit([&run, r] { run(r); });
      }
      for (int r = 0; r < g; ++r) h->workers[r]->Wait();
    }
    h->last_launches = 0;
    for (int r = 0; r < g; ++r) {
      if (st[r] != cudaSuccess) return Fail(h, RN_ERR_CUDA, h->replicas[r]->error());
      h->last_launches += h->replicas[r]->last_launches();
    }
    double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    if (h->lat_ms.size() < (1u << 20)) h->lat_ms.push_back(ms);
    h->calls += 1;
    h->images += n;
    return RN_OK;
  });
}

int rn_decode_jpeg_u8_bgr(rn_handle* h, const uint8_t* file, uint64_t size, uint8_t* out, uint64_t out_capacity,
                          int32_t* height, int32_t* width, int32_t* status) {
  return Guarded(h, [&]() -> int {
    if (!h) return RN_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> lock(h->mu);
    if (!file || !height || !width || !status) return Fail(h, RN_ERR_INVALID_ARG, "null argument");
    if (h->replicas.empty()) return Fail(h, RN_ERR_CUDA, "host-only handle (n_devices = 0): the decoder's second half runs on the device");
    int hh = 0, ww = 0;
    cudaError_t e = h->replicas[0]->DecodeJpeg(file, size, out, out_capacity, &hh, &ww, status);
    if (e != cudaSuccess) return Fail(h, e == cudaErrorInvalidValue ? RN_ERR_INVALID_ARG : RN_ERR_CUDA, h->replicas[0]->error());
    *height = hh;
    *width = ww;
    return RN_OK;
  });
}

int rn_jpeg_info(const uint8_t* file, uint64_t size, int64_t info[8]) {
  if (!file || !info) return RN_ERR_INVALID_ARG;
  try {
    rn::JpegInfo f;
    const int st = rn::JpegParseHeader(file, size, &f);
    const bool swap = f.orientation >= 5;
    info[0] = st;
    info[1] = swap ? f.height : f.width;
    info[2] = swap ? f.width : f.height;
    info[3] = f.ncomp;
    info[4] = f.hmax;
    info[5] = f.vmax;
    info[6] = f.orientation;
    info[7] = static_cast<int64_t>(f.coef_count);
    return RN_OK;
  } catch (...) {
    return RN_ERR_INTERNAL;
  }
}

int rn_jpeg_prepare_scan(const uint8_t* file, uint64_t size, uint8_t* stream, uint64_t capacity, int32_t* sub_seg,
                         int64_t info[4]) {
  if (!file || !stream || !sub_seg || !info) return RN_JPEG_CORRUPT;
  try {
    rn::JpegInfo f;
    int st = rn::JpegParseHeader(file, size, &f);
    if (st != rn::kJpegOk) return st;
    if (capacity < rn::JpegStreamCapacity(f, size)) return RN_JPEG_CORRUPT;
    auto plan = std::make_unique<rn::JpegScanPlan>();
    st = rn::JpegPrepareScan(file, size, f, plan.get(), stream, capacity, sub_seg);
    if (st != rn::kJpegOk) return st;
    info[0] = static_cast<int64_t>(plan->stream_bytes);
    info[1] = plan->n_seg;
    info[2] = plan->bpm;
    info[3] = plan->total_blocks;
    return RN_JPEG_OK;
  } catch (...) {
    return RN_JPEG_CORRUPT;
  }
}

int rn_jpeg_coefficients(const uint8_t* file, uint64_t size, int16_t* coefs, uint64_t capacity) {
  if (!file || !coefs) return RN_JPEG_CORRUPT;
  try {
    rn::JpegInfo f;
    const int st = rn::JpegParseHeader(file, size, &f);
    if (st != rn::kJpegOk) return st;
    if (capacity < f.coef_count) return RN_JPEG_CORRUPT;
    return rn::JpegDecodeCoefficients(file, size, f, coefs);
  } catch (...) {
    return RN_JPEG_CORRUPT;
  }
}

int rn_infer_yuv420(rn_handle* h, const uint8_t* y, const uint8_t* u, const uint8_t* v, int32_t y_size, int32_t u_size,
                    int32_t v_size, int32_t width, int32_t height, int32_t y_row_stride, int32_t uv_row_stride,
                    int32_t uv_pixel_stride, int32_t rotation, int64_t* top1, float* probs, float* logits,
                    uint8_t* rgb_out) {
  return Guarded(h, [&]() -> int {
    if (!h) return RN_ERR_INVALID_ARG;
    std::lock_guard<std::mutex> lock(h->mu);
    if (!y || !u || !v || width < 2 || height < 2 || y_row_stride < width || uv_row_stride < 1 || uv_pixel_stride < 1)
      return Fail(h, RN_ERR_INVALID_ARG, "null plane or bad frame geometry");
    rotation %= 360;
    if (rotation < 0) rotation += 360;
    if (rotation % 90 != 0) return Fail(h, RN_ERR_INVALID_ARG, "rotation must be a multiple of 90 degrees");
    // every (row, column) the kernel can touch must lie inside the planes the caller handed over
    const long long y_need = static_cast<long long>(y_row_stride) * (height - 1) + width;
    const long long uv_need = static_cast<long long>(uv_row_stride) * ((height - 1) >> 1) +
                              static_cast<long long>((width - 1) >> 1) * uv_pixel_stride + 1;
    if (y_size < y_need || u_size < uv_need || v_size < uv_need)
      return Fail(h, RN_ERR_INVALID_ARG, "plane buffers are smaller than the frame geometry requires");
    if (h->replicas.empty()) return Fail(h, RN_ERR_CUDA, "host-only handle (n_devices = 0): there is no CPU inference path");
    if (!h->loaded) return Fail(h, RN_ERR_NOT_LOADED, "weights have not been loaded (rn_load_tf_checkpoint)");
    auto t0 = std::chrono::steady_clock::now();
    rn::Replica* r = h->replicas[0].get();
    rn::YuvFrame f{width, height, y_row_stride, uv_row_stride, uv_pixel_stride, rotation};
    if (r->InferYuv420(y, u, v, y_size, u_size, v_size, f, top1, probs, logits, rgb_out) != cudaSuccess)
      return Fail(h, RN_ERR_CUDA, r->error());
    h->last_launches = r->last_launches();
    double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    if (h->lat_ms.size() < (1u << 20)) h->lat_ms.push_back(ms);
    h->calls += 1;
    h->images += 1;
    return RN_OK;
  });
}

int rn_center_crop_rect(int32_t hgt, int32_t wid, int32_t* y0, int32_t* x0, int32_t* side) {
  if (hgt <= 0 || wid <= 0 || !y0 || !x0 || !side) return RN_ERR_INVALID_ARG;
  // reference network.py:139: offset = abs((w - h) // 2) with Python floor division
  int d = wid - hgt;
  int fl = d >= 0 ? d / 2 : -((-d + 1) / 2);
  int off = fl < 0 ? -fl : fl;
  *y0 = 0;
  *x0 = 0;
  if (hgt < wid) {
    *x0 = off;
    *side = hgt;
  } else if (wid < hgt) {
    *y0 = off;
    *side = wid;
  } else {
    *side = hgt;
  }
  return RN_OK;
}

int rn_flat_len(const rn_handle* h) { return h ? h->shape.flat_len : -1; }
int rn_num_kernel_launches(const rn_handle* h) { return h ? h->last_launches : -1; }

int rn_get_folded(rn_handle* h, const char* name, float* out, int64_t capacity, int64_t* size) {
  if (!h || !name || !size) return RN_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lock(h->mu);
  if (!h->loaded) return Fail(h, RN_ERR_NOT_LOADED, "weights have not been loaded");
  const std::vector<double>* v = nullptr;
  std::string nm(name);
  auto layer_of = [&](const std::string& prefix, int limit, int* idx) {
    if (nm.compare(0, prefix.size(), prefix) != 0) return false;
    size_t pos = prefix.size();
    int i = 0;
    bool any = false;
    while (pos < nm.size() && nm[pos] >= '0' && nm[pos] <= '9') {
      i = i * 10 + (nm[pos++] - '0');
      any = true;
    }
    if (!any || i >= limit || pos >= nm.size() || nm[pos] != '/') return false;
    *idx = i;
    nm = nm.substr(pos + 1);
    return true;
  };
  int i = 0;
  const rn::FoldedNet& f = h->folded;
  if (nm == "conv0_u8bgr/w") v = &f.conv0_u8bgr.w;
  else if (nm == "conv0_u8bgr/b") v = &f.conv0_u8bgr.b;
  else if (nm == "conv0_u8rgb/w") v = &f.conv0_u8rgb.w;
  else if (nm == "conv0_u8rgb/b") v = &f.conv0_u8rgb.b;
  else if (nm == "conv0_f32rgb/w") v = &f.conv0_f32rgb.w;
  else if (nm == "conv0_f32rgb/b") v = &f.conv0_f32rgb.b;
  else if (layer_of("conv", rn::kNumConvs, &i)) v = nm == "w" ? &f.conv[i].w : nm == "b" ? &f.conv[i].b : nullptr;
  else if (layer_of("join", rn::kNumConvs, &i)) v = nm == "a" ? &f.join[i].a : nm == "b" ? &f.join[i].b : nm == "c" ? &f.join[i].c : nullptr;
  else if (layer_of("dense", rn::kNumDense, &i)) v = nm == "w" ? &f.dense[i].w : nm == "b" ? &f.dense[i].b : nullptr;
  if (!v || v->empty()) return Fail(h, RN_ERR_INVALID_ARG, std::string("unknown folded tensor '") + name + "'");
  *size = static_cast<int64_t>(v->size());
  if (out) {
    if (capacity < *size) return Fail(h, RN_ERR_INVALID_ARG, "output buffer too small");
    for (size_t k = 0; k < v->size(); ++k) out[k] = static_cast<float>((*v)[k]);
  }
  return RN_OK;
}

int rn_debug_activation(rn_handle* h, int32_t layer, float* out, int64_t capacity, int64_t* size, int32_t dims[4]) {
  if (!h || !size || !dims) return RN_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lock(h->mu);
  std::vector<float> v;
  int d[4];
  if (h->replicas.empty()) return Fail(h, RN_ERR_CUDA, "host-only handle has no activations");
  if (h->replicas[0]->DebugActivation(layer, &v, d) != cudaSuccess) return Fail(h, RN_ERR_CUDA, h->replicas[0]->error());
  for (int k = 0; k < 4; ++k) dims[k] = d[k];
  *size = static_cast<int64_t>(v.size());
  if (out) {
    if (capacity < *size) return Fail(h, RN_ERR_INVALID_ARG, "output buffer too small");
    std::memcpy(out, v.data(), v.size() * sizeof(float));
  }
  return RN_OK;
}

int rn_get_stats(rn_handle* h, double* p50_ms, double* p99_ms, int64_t* calls, int64_t* images) {
  if (!h) return RN_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lock(h->mu);
  std::vector<double> v = h->lat_ms;
  std::sort(v.begin(), v.end());
  auto pct = [&](double p) { return v.empty() ? 0.0 : v[std::min(v.size() - 1, static_cast<size_t>(p * v.size()))]; };
  if (p50_ms) *p50_ms = pct(0.50);
  if (p99_ms) *p99_ms = pct(0.99);
  if (calls) *calls = h->calls;
  if (images) *images = h->images;
  return RN_OK;
}

int rn_get_jpeg_counters(rn_handle* h, int64_t* device_huffman_files, int64_t* host_huffman_files) {
  if (!h) return RN_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lock(h->mu);
  long long d = 0, c = 0;
  for (auto& r : h->replicas) {
    long long rd = 0, rc = 0;
    r->jpeg_counters(&rd, &rc);
    d += rd;
    c += rc;
  }
  if (device_huffman_files) *device_huffman_files = d;
  if (host_huffman_files) *host_huffman_files = c;
  return RN_OK;
}

int rn_set_profiling(rn_handle* h, int32_t enabled) {
  if (!h) return RN_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lock(h->mu);
  for (auto& r : h->replicas) r->set_profiling(enabled != 0);
  return RN_OK;
}

int rn_get_profile(rn_handle* h, int32_t capacity, int32_t* count, char names[][32], double* ms, int32_t* launches) {
  if (!h || !count) return RN_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lock(h->mu);
  if (h->replicas.empty()) return Fail(h, RN_ERR_CUDA, "host-only handle");
  std::vector<rn::Replica::KernelTime> v;
  if (h->replicas[0]->ProfileResults(&v) != cudaSuccess) return Fail(h, RN_ERR_CUDA, h->replicas[0]->error());
  *count = static_cast<int32_t>(v.size());
  for (int i = 0; i < *count && i < capacity; ++i) {
    if (names) {
      std::strncpy(names[i], v[i].name.c_str(), 31);
      names[i][31] = 0;
    }
    if (ms) ms[i] = v[i].ms;
    if (launches) launches[i] = v[i].launches;
  }
  return RN_OK;
}

int rn_reset_stats(rn_handle* h) {
  if (!h) return RN_ERR_INVALID_ARG;
  std::lock_guard<std::mutex> lock(h->mu);
  h->lat_ms.clear();
  h->calls = h->images = 0;
  return RN_OK;
}

}  // extern "C"
