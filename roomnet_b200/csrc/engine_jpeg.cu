// Replica entry points of the JPEG front end: cv2.imread + infer_optimized of the reference's directory loop
// (infer.py:79-82) for files that are baseline JPEGs.  Entropy decoding on a few host threads (jpeg_host.cpp), then on
// the device: inverse DCT -> upsampling + colour conversion (kernels_jpeg.cu) -> centre crop + cv2-identical resize of
// the whole micro-batch -> forward pass.  The decoded photo never exists in host memory.
#include <algorithm>
#include <atomic>
#include <cstring>
#include <thread>

#include "engine.h"
#include "jpeg_host.h"

namespace rn {

#define RN_CUDA(expr)                                                                                   \
  do {                                                                                                  \
    cudaError_t _e = (expr);                                                                            \
    if (_e != cudaSuccess) {                                                                            \
      err_ = std::string(#expr) + ": " + cudaGetErrorString(_e) + " (device " + std::to_string(device_) + ")"; \
      return _e;                                                                                        \
    }                                                                                                   \
  } while (0)

namespace {
inline size_t Align(size_t v, size_t a) { return (v + a - 1) / a * a; }

// reference network.py:137-146: offset = abs((w - h) // 2) with Python floor division
CropDesc CentreCrop(int h, int w, size_t offset) {
  const int d = w - h;
  const int fl = d >= 0 ? d / 2 : -((-d + 1) / 2);
  const int o = fl < 0 ? -fl : fl;
  CropDesc c{};
  c.offset = offset;
  c.W = w;
  c.side = std::min(h, w);
  c.cy = h > w ? o : 0;
  c.cx = w > h ? o : 0;
  return c;
}
}  // namespace

cudaError_t Replica::GrowJpegBuffers(size_t coef_bytes, size_t sample_bytes, size_t raw_bytes, int n_images) {
  auto grow = [&](void** p, size_t* cap, size_t need, bool host) -> cudaError_t {
    if (need <= *cap) return cudaSuccess;
    cudaError_t e = cudaStreamSynchronize(compute_);
    if (e != cudaSuccess) return e;
    if (*p) (host ? cudaFreeHost(*p) : cudaFree(*p));
    *p = nullptr;
    *cap = 0;
    const size_t want = need + need / 4;  // headroom: photo sizes vary from batch to batch
    e = host ? cudaMallocHost(p, want) : cudaMalloc(p, want);
    if (e == cudaSuccess) *cap = want;
    return e;
  };
  RN_CUDA(grow(reinterpret_cast<void**>(&h_coef_), &h_coef_cap_, coef_bytes, true));
  RN_CUDA(grow(reinterpret_cast<void**>(&d_coef_), &d_coef_cap_, coef_bytes, false));
  RN_CUDA(grow(reinterpret_cast<void**>(&d_samples_), &d_samples_cap_, sample_bytes, false));
  RN_CUDA(grow(reinterpret_cast<void**>(&d_raw_), &d_raw_cap_, raw_bytes, false));
  const size_t meta = static_cast<size_t>(n_images) * (3 * sizeof(JpegPlaneDesc) + sizeof(JpegImageDesc) + 3 * 64 * sizeof(uint16_t));
  RN_CUDA(grow(reinterpret_cast<void**>(&d_jmeta_), &d_jmeta_cap_, meta, false));
  return cudaSuccess;
}

// Decodes the listed files into oriented BGR images in d_raw_ (enqueued on compute_, not synchronised).
// crops[k] / ok[k] describe list entry k; status (optional) receives the per-file JpegStatus at index[k].
cudaError_t Replica::JpegToRaw(const uint8_t* const* files, const size_t* sizes, const std::vector<int>& index,
                               const std::vector<JpegInfo>& info, int threads, std::vector<CropDesc>* crops,
                               std::vector<char>* ok, int32_t* status) {
  const int m = static_cast<int>(index.size());
  std::vector<size_t> coef_off(m), raw_off(m);
  std::vector<size_t> plane_off(static_cast<size_t>(m) * 3, 0);
  size_t coef_total = 0, sample_total = 0, raw_total = 0;
  for (int k = 0; k < m; ++k) {
    const JpegInfo& f = info[k];
    coef_off[k] = coef_total;
    coef_total += Align(f.coef_count, 64);
    for (int c = 0; c < f.ncomp; ++c) {
      plane_off[3 * k + c] = sample_total;
      sample_total += Align(static_cast<size_t>(f.comp[c].wblocks) * f.comp[c].hblocks * 64, 256);
    }
    raw_off[k] = raw_total;
    raw_total += Align(static_cast<size_t>(f.width) * f.height * 3, 256);
  }
  cudaError_t e = GrowJpegBuffers(coef_total * sizeof(int16_t), sample_total, raw_total, m);
  if (e != cudaSuccess) return e;
  // the pinned coefficient buffer is reused from batch to batch: the previous upload must have left it
  RN_CUDA(cudaStreamSynchronize(compute_));

  // ---- entropy decoding: files are independent, a handful of host threads share them ----
  std::vector<int> st(m, kJpegOk);
  {
    std::atomic<int> next{0};
    auto work = [&]() {
      for (int k = next.fetch_add(1); k < m; k = next.fetch_add(1))
        st[k] = JpegDecodeCoefficients(files[index[k]], sizes[index[k]], info[k], h_coef_ + coef_off[k]);
    };
    const int nt = std::max(1, std::min(threads, m));
    std::vector<std::thread> pool;
    for (int t = 1; t < nt; ++t) pool.emplace_back(work);
    work();
    for (auto& t : pool) t.join();
  }

  // ---- descriptors of what decoded cleanly ----
  std::vector<JpegPlaneDesc> planes;
  std::vector<JpegImageDesc> images;
  std::vector<uint16_t> quant;
  crops->assign(m, CropDesc{});
  ok->assign(m, 0);
  for (int k = 0; k < m; ++k) {
    if (status) status[index[k]] = st[k];
    if (st[k] != kJpegOk) continue;
    const JpegInfo& f = info[k];
    JpegImageDesc im{};
    for (int c = 0; c < f.ncomp; ++c) {
      JpegPlaneDesc p{};
      p.coef_offset = coef_off[k] + f.comp[c].coef_offset;
      p.plane_offset = plane_off[3 * k + c];
      p.wblocks = f.comp[c].wblocks;
      p.hblocks = f.comp[c].hblocks;
      p.quant_index = static_cast<int>(quant.size() / 64);
      quant.insert(quant.end(), f.quant[f.comp[c].tq], f.quant[f.comp[c].tq] + 64);
      planes.push_back(p);
      im.plane[c] = p.plane_offset;
      im.pitch[c] = p.wblocks * 8;
    }
    im.width = f.width;
    im.height = f.height;
    im.ncomp = f.ncomp;
    im.hs = f.hmax;
    im.vs = f.vmax;
    im.cdw = f.ncomp == 3 ? f.comp[1].dw : f.width;
    im.cdh = f.ncomp == 3 ? f.comp[1].dh : f.height;
    im.orientation = f.orientation;
    const bool swap = f.orientation >= 5;
    const int ow = swap ? f.height : f.width, oh = swap ? f.width : f.height;
    im.out_w = ow;
    im.out_offset = raw_off[k];
    images.push_back(im);
    (*crops)[k] = CentreCrop(oh, ow, raw_off[k]);
    (*ok)[k] = 1;
  }
  if (images.empty()) return cudaSuccess;
  char* meta = static_cast<char*>(d_jmeta_);
  JpegPlaneDesc* d_planes = reinterpret_cast<JpegPlaneDesc*>(meta);
  JpegImageDesc* d_images = reinterpret_cast<JpegImageDesc*>(meta + static_cast<size_t>(m) * 3 * sizeof(JpegPlaneDesc));
  uint16_t* d_quant = reinterpret_cast<uint16_t*>(meta + static_cast<size_t>(m) * (3 * sizeof(JpegPlaneDesc) + sizeof(JpegImageDesc)));
  RN_CUDA(cudaMemcpyAsync(d_coef_, h_coef_, coef_total * sizeof(int16_t), cudaMemcpyHostToDevice, compute_));
  RN_CUDA(cudaMemcpyAsync(d_planes, planes.data(), planes.size() * sizeof(JpegPlaneDesc), cudaMemcpyHostToDevice, compute_));
  RN_CUDA(cudaMemcpyAsync(d_images, images.data(), images.size() * sizeof(JpegImageDesc), cudaMemcpyHostToDevice, compute_));
  RN_CUDA(cudaMemcpyAsync(d_quant, quant.data(), quant.size() * sizeof(uint16_t), cudaMemcpyHostToDevice, compute_));
  RN_CUDA(JpegIdct(d_coef_, d_planes, static_cast<int>(planes.size()), d_quant, d_samples_, compute_));
  RN_CUDA(JpegColor(d_samples_, d_images, static_cast<int>(images.size()), d_raw_, compute_));
  last_launches_ += 2;
  // planes / images / quant are pageable vectors: their copies were staged by the runtime before the calls returned
  return cudaSuccess;
}

cudaError_t Replica::DecodeJpeg(const uint8_t* file, size_t size, uint8_t* out, size_t capacity, int* height, int* width,
                                int32_t* status) {
  {
    cudaError_t ew = WaitHost(~0ull);
    if (ew != cudaSuccess) return ew;
  }
  RN_CUDA(cudaSetDevice(device_));
  std::vector<JpegInfo> info(1);
  const JpegStatus hs = JpegParseHeader(file, size, &info[0]);
  *status = hs;
  if (hs != kJpegOk) return cudaSuccess;
  const bool swap = info[0].orientation >= 5;
  *height = swap ? info[0].width : info[0].height;
  *width = swap ? info[0].height : info[0].width;
  const size_t bytes = static_cast<size_t>(info[0].width) * info[0].height * 3;
  if (!out) return cudaSuccess;
  if (capacity < bytes) {
    err_ = "output buffer smaller than height * width * 3";
    return cudaErrorInvalidValue;
  }
  std::vector<CropDesc> crops;
  std::vector<char> ok;
  last_launches_ = 0;
  const size_t sz = size;
  cudaError_t e = JpegToRaw(&file, &sz, std::vector<int>{0}, info, 1, &crops, &ok, status);
  if (e != cudaSuccess) return e;
  if (!ok[0]) return cudaSuccess;
  RN_CUDA(cudaMemcpyAsync(out, d_raw_ + crops[0].offset, bytes, cudaMemcpyDeviceToHost, compute_));
  RN_CUDA(cudaStreamSynchronize(compute_));
  return cudaSuccess;
}

cudaError_t Replica::InferJpegs(const uint8_t* const* files, const size_t* sizes, int n, int threads, int64_t* top1,
                                float* probs, float* logits, int32_t* status) {
  {
    cudaError_t ew = WaitHost(~0ull);  // uses staging slot 0 and activation set 0 on the replica's own stream
    if (ew != cudaSuccess) return ew;
  }
  RN_CUDA(cudaSetDevice(device_));
  last_launches_ = 0;
  const int S = shape_.im_side, C = shape_.num_classes;
  if (!d_descs_) RN_CUDA(Alloc(&d_descs_, static_cast<size_t>(max_batch_) * sizeof(CropDesc)));
  // headers first: geometry decides the micro-batches (at most max_batch files and ~1.5 GB of device staging each)
  std::vector<JpegInfo> all(n);
  for (int i = 0; i < n; ++i) status[i] = JpegParseHeader(files[i], sizes[i], &all[i]);
  constexpr size_t kStagingCap = size_t{3} << 29;
  int i = 0;
  std::vector<int> index;
  std::vector<JpegInfo> info;
  std::vector<CropDesc> crops, packed;
  std::vector<char> ok;
  std::vector<int> where;
  while (i < n) {
    index.clear();
    info.clear();
    size_t staged = 0;
    for (; i < n && static_cast<int>(index.size()) < max_batch_; ++i) {
      if (status[i] != kJpegOk) continue;
      const size_t px = static_cast<size_t>(all[i].width) * all[i].height;
      const size_t need = all[i].coef_count * 2 + all[i].coef_count + px * 3 + 4096;
      if (!index.empty() && staged + need > kStagingCap) break;
      staged += need;
      index.push_back(i);
      info.push_back(all[i]);
    }
    if (index.empty()) continue;
    cudaError_t e = JpegToRaw(files, sizes, index, info, threads, &crops, &ok, status);
    if (e != cudaSuccess) return e;
    packed.clear();
    where.clear();
    for (size_t k = 0; k < index.size(); ++k) {
      if (!ok[k]) continue;
      packed.push_back(crops[k]);
      where.push_back(index[k]);
    }
    const int m = static_cast<int>(packed.size());
    if (m == 0) continue;
    RN_CUDA(cudaMemcpyAsync(d_descs_, packed.data(), m * sizeof(CropDesc), cudaMemcpyHostToDevice, compute_));
    RN_CUDA(CropResizeBatchU8(d_raw_, static_cast<const CropDesc*>(d_descs_), m, static_cast<uint8_t*>(d_in_[0]), S,
                              compute_));
    ++last_launches_;
    cur_ = &sets_[0];
    const HostOut ho = Out(0);
    e = ForwardDevice(d_in_[0], InputKind::kU8Bgr, m, ho.top1, ho.probs, ho.logits, compute_);
    if (e != cudaSuccess) return e;
    RN_CUDA(cudaStreamSynchronize(compute_));
    for (int k = 0; k < m; ++k) {  // scatter: files that fell out keep their slots untouched
      const int dst = where[k];
      if (top1) top1[dst] = ho.top1[k];
      if (probs) std::memcpy(probs + static_cast<size_t>(dst) * C, ho.probs + static_cast<size_t>(k) * C, C * sizeof(float));
      if (logits) std::memcpy(logits + static_cast<size_t>(dst) * C, ho.logits + static_cast<size_t>(k) * C, C * sizeof(float));
    }
  }
  return cudaSuccess;
}

}  // namespace rn
