// Replica entry points of the JPEG front end: cv2.imread + infer_optimized of the reference's directory loop
// (infer.py:79-82) for files that are baseline JPEGs.  Host threads strip the byte stuffing of the files into a ring of
// pinned buffers (jpeg_host.cpp); on the device: Huffman decoding (kernels_jpeg_huff.cu) -> inverse DCT -> upsampling +
// colour conversion (kernels_jpeg.cu) -> centre crop + cv2-identical resize of the whole micro-batch -> forward pass.
// The decoded photo never exists in host memory.  Files the device Huffman stage does not take (several scans, damaged
// streams) go through the host Huffman decoder and the same device kernels.
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstring>
#include <memory>
#include <mutex>
#include <thread>

#include <nvtx3/nvToolsExt.h>

#include "engine.h"
#include "jpeg_host.h"
#include "roomnet.h"

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
struct NvtxRangeJpeg {
  explicit NvtxRangeJpeg(const char* name) { nvtxRangePushA(name); }
  ~NvtxRangeJpeg() { nvtxRangePop(); }
};
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

// One micro-batch of files on its way through the decoder.
struct Replica::JpegBatch {
  std::vector<int> index;      // positions in the caller's list
  std::vector<JpegInfo> info;
  std::vector<size_t> coef_off;  // int16 units into the pinned coefficient buffer
  size_t coef_total = 0;
  std::vector<int> st;         // JpegStatus per file after entropy decoding
  // device Huffman stage: layout of the pinned staging buffer [streams | sub_seg] and the per-file scan plans
  std::vector<size_t> stream_off;
  size_t stream_total = 0, stage_bytes = 0;
  std::vector<JpegScanPlan> plans;
  void LayOutStreams(const size_t* sizes) {
    const size_t m = index.size();
    stream_off.resize(m);
    stream_total = 0;
    for (size_t k = 0; k < m; ++k) {
      stream_off[k] = stream_total;
      stream_total += JpegStreamCapacity(info[k], sizes[index[k]]);
    }
    stream_total += kSubseqBytes;  // look-ahead slack behind the last stream
    stage_bytes = stream_total + stream_total / kSubseqBytes * sizeof(int32_t);
    plans.resize(m);
    st.assign(m, kJpegUnsupported);
  }
  // host half for entry k: strip the byte stuffing into the staging buffer
  void PrepareStream(const uint8_t* const* files, const size_t* sizes, size_t k, uint8_t* h_stage) {
    const size_t cap = (k + 1 < index.size() ? stream_off[k + 1] : stream_total - kSubseqBytes) - stream_off[k];
    int32_t* h_sub_seg = reinterpret_cast<int32_t*>(h_stage + stream_total);
    st[k] = JpegPrepareScan(files[index[k]], sizes[index[k]], info[k], &plans[k], h_stage + stream_off[k], cap,
                            h_sub_seg + stream_off[k] / kSubseqBytes);
  }
};

cudaError_t Replica::GrowJpegBuffers(size_t coef_bytes, size_t sample_bytes, size_t raw_bytes, int n_images, int n_host) {
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
  for (int b = 0; b < n_host; ++b)
    RN_CUDA(grow(reinterpret_cast<void**>(&h_coef_[b]), &h_coef_cap_[b], coef_bytes, true));
  RN_CUDA(grow(reinterpret_cast<void**>(&d_coef_), &d_coef_cap_, coef_bytes, false));
  RN_CUDA(grow(reinterpret_cast<void**>(&d_samples_), &d_samples_cap_, sample_bytes, false));
  RN_CUDA(grow(reinterpret_cast<void**>(&d_raw_), &d_raw_cap_, raw_bytes, false));
  const size_t meta = static_cast<size_t>(n_images) * (3 * sizeof(JpegPlaneDesc) + sizeof(JpegImageDesc) + 3 * 64 * sizeof(uint16_t));
  RN_CUDA(grow(reinterpret_cast<void**>(&d_jmeta_), &d_jmeta_cap_, meta, false));
  return cudaSuccess;
}

// Host half of a batch: entropy decoding into a pinned buffer; files are independent, a handful of threads share them.
void Replica::JpegDecodeHost(const uint8_t* const* files, const size_t* sizes, JpegBatch* b, int16_t* h_coef, int threads) {
  const int m = static_cast<int>(b->index.size());
  b->st.assign(m, kJpegOk);
  std::atomic<int> next{0};
  auto work = [&]() {
    for (int k = next.fetch_add(1); k < m; k = next.fetch_add(1))
      b->st[k] = JpegDecodeCoefficients(files[b->index[k]], sizes[b->index[k]], b->info[k], h_coef + b->coef_off[k]);
  };
  const int nt = std::max(1, std::min(threads, m));
  std::vector<std::thread> pool;
  for (int t = 1; t < nt; ++t) pool.emplace_back(work);
  work();
  for (auto& t : pool) t.join();
}

// Device half: coefficients -> oriented BGR images in d_raw_ (enqueued on compute_, not synchronised).
// crops[k] / ok[k] describe entry k of the batch; status (optional) receives the per-file JpegStatus.
cudaError_t Replica::JpegToRaw(const JpegBatch& b, const int16_t* h_coef, std::vector<CropDesc>* crops,
                               std::vector<char>* ok, int32_t* status) {
  const int m = static_cast<int>(b.index.size());
  std::vector<size_t> raw_off(m);
  std::vector<size_t> plane_off(static_cast<size_t>(m) * 3, 0);
  size_t sample_total = 0, raw_total = 0;
  for (int k = 0; k < m; ++k) {
    const JpegInfo& f = b.info[k];
    for (int c = 0; c < f.ncomp; ++c) {
      plane_off[3 * k + c] = sample_total;
      sample_total += Align(static_cast<size_t>(f.comp[c].wblocks) * f.comp[c].hblocks * 64, 256);
    }
    raw_off[k] = raw_total;
    {
      const bool swap = f.orientation >= 5;
      const size_t ow = swap ? f.height : f.width, oh = swap ? f.width : f.height;
      raw_total += Align(Align(ow, 4) * oh * 3, 256);
    }
  }
  cudaError_t e = GrowJpegBuffers(b.coef_total * sizeof(int16_t), sample_total, raw_total, m, 0);
  if (e != cudaSuccess) return e;
  std::vector<JpegPlaneDesc> planes;
  std::vector<JpegImageDesc> images;
  std::vector<uint16_t> quant;
  crops->assign(m, CropDesc{});
  ok->assign(m, 0);
  for (int k = 0; k < m; ++k) {
    if (status) status[b.index[k]] = b.st[k];
    if (b.st[k] != kJpegOk) continue;
    const JpegInfo& f = b.info[k];
    JpegImageDesc im{};
    for (int c = 0; c < f.ncomp; ++c) {
      JpegPlaneDesc p{};
      p.coef_offset = b.coef_off[k] + f.comp[c].coef_offset;
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
    im.out_pitch = static_cast<int>(Align(ow, 4));
    im.out_offset = raw_off[k];
    images.push_back(im);
    (*crops)[k] = CentreCrop(oh, ow, raw_off[k]);
    (*crops)[k].W = im.out_pitch;  // the crop kernel only uses it as the row pitch
    (*ok)[k] = 1;
  }
  if (images.empty()) return cudaSuccess;
  char* meta = static_cast<char*>(d_jmeta_);
  JpegPlaneDesc* d_planes = reinterpret_cast<JpegPlaneDesc*>(meta);
  JpegImageDesc* d_images = reinterpret_cast<JpegImageDesc*>(meta + static_cast<size_t>(m) * 3 * sizeof(JpegPlaneDesc));
  uint16_t* d_quant = reinterpret_cast<uint16_t*>(meta + static_cast<size_t>(m) * (3 * sizeof(JpegPlaneDesc) + sizeof(JpegImageDesc)));
  if (h_coef)  // (device Huffman decoding: the coefficients are already in d_coef_)
    RN_CUDA(cudaMemcpyAsync(d_coef_, h_coef, b.coef_total * sizeof(int16_t), cudaMemcpyHostToDevice, compute_));
  // the descriptor vectors are pageable: the runtime stages such copies before the call returns
  RN_CUDA(cudaMemcpyAsync(d_planes, planes.data(), planes.size() * sizeof(JpegPlaneDesc), cudaMemcpyHostToDevice, compute_));
  RN_CUDA(cudaMemcpyAsync(d_images, images.data(), images.size() * sizeof(JpegImageDesc), cudaMemcpyHostToDevice, compute_));
  RN_CUDA(cudaMemcpyAsync(d_quant, quant.data(), quant.size() * sizeof(uint16_t), cudaMemcpyHostToDevice, compute_));
  RN_CUDA(JpegIdct(d_coef_, d_planes, static_cast<int>(planes.size()), d_quant, d_samples_, compute_));
  RN_CUDA(JpegColor(d_samples_, d_images, static_cast<int>(images.size()), d_raw_, compute_));
  last_launches_ += 2;
  return cudaSuccess;
}

// Device-Huffman stage of one batch: host threads strip the byte stuffing of each file's scan into a pinned buffer
// (here, or ahead of time by the caller's pipeline: h_prepared), the device decodes the streams into d_coef_
// (enqueued on compute_).  b->st[k] = kJpegOk for the files handed to the
// device, anything else = leave that file to the host decoder.  (*err)[k] becomes valid after the stream has been
// synchronised: non-zero = the device found the stream damaged.
cudaError_t Replica::JpegHuffUpload(const uint8_t* const* files, const size_t* sizes, JpegBatch* b, int threads,
                                    const uint8_t* h_prepared, int slot) {
  const int m = static_cast<int>(b->index.size());
  HuffStage& stg = huff_stage_[slot];
  stg.any = false;
  stg.file_of.clear();
  const uint8_t* h_stream = h_prepared;
  if (!h_prepared) {  // not staged by the caller's pipeline: do the host half here
    b->LayOutStreams(sizes);
    cudaError_t e = GrowJpegBuffers((b->stage_bytes + 1) / 2 * 2, 0, 0, 0, 1);  // h_coef_[0] doubles as the staging buffer
    if (e != cudaSuccess) return e;
    RN_CUDA(cudaStreamSynchronize(compute_));  // an earlier upload may still be reading the pinned buffer
    RN_CUDA(cudaStreamSynchronize(copy_));
    uint8_t* stage = reinterpret_cast<uint8_t*>(h_coef_[0]);
    std::atomic<int> next{0};
    auto work = [&]() {
      for (int k = next.fetch_add(1); k < m; k = next.fetch_add(1)) b->PrepareStream(files, sizes, k, stage);
    };
    const int nt = std::max(1, std::min(threads, m));
    std::vector<std::thread> pool;
    for (int t = 1; t < nt; ++t) pool.emplace_back(work);
    work();
    for (auto& t : pool) t.join();
    h_stream = stage;
  }
  const std::vector<size_t>& stream_off = b->stream_off;
  const std::vector<JpegScanPlan>& plans = b->plans;
  const size_t stream_total = b->stream_total, stage_bytes = b->stage_bytes;
  const size_t n_sub_total = stream_total / kSubseqBytes;
  // ---- descriptors ----
  std::vector<HuffFileDesc> fds;
  std::vector<HuffBlockDesc> bds;
  std::vector<DevHuffTable> tabs;
  std::vector<int> file_of;  // batch entry of each descriptor
  size_t dc_total = 0;
  for (int k = 0; k < m; ++k) {
    if (b->st[k] != kJpegOk) continue;
    const JpegInfo& f = b->info[k];
    const JpegScanPlan& p = plans[k];
    HuffFileDesc d{};
    d.stream_off = stream_off[k];
    d.sub_base = static_cast<unsigned>(stream_off[k] / kSubseqBytes);
    d.n_sub = static_cast<unsigned>(p.stream_bytes / kSubseqBytes);
    d.total_blocks = p.total_blocks;
    d.seg_blocks = p.seg_blocks;
    d.bpm = p.bpm;
    d.mcus_x = p.mcus_x;
    d.ncomp = f.ncomp;
    d.table_index = static_cast<int>(tabs.size());
    d.dc_off = dc_total;
    dc_total += Align(p.total_blocks, 64);
    for (int c = 0; c < f.ncomp; ++c) {
      d.coef_off[c] = b->coef_off[k] + f.comp[c].coef_offset;
      d.wblocks[c] = f.comp[c].wblocks;
      d.hblocks[c] = f.comp[c].hblocks;
      d.comp_h[c] = f.comp[c].h;
      d.comp_v[c] = f.comp[c].v;
    }
    std::memcpy(d.blk_comp, p.blk_comp, 8);
    std::memcpy(d.blk_hh, p.blk_hh, 8);
    std::memcpy(d.blk_vv, p.blk_vv, 8);
    tabs.insert(tabs.end(), p.tab, p.tab + 6);
    for (unsigned s0 = 0; s0 < d.n_sub; s0 += 256) bds.push_back(HuffBlockDesc{static_cast<int>(fds.size()), s0});
    fds.push_back(d);
    file_of.push_back(k);
  }
  if (fds.empty()) return cudaSuccess;
  const int nf = static_cast<int>(fds.size()), nb = static_cast<int>(bds.size());
  stg.any = true;
  stg.nf = nf;
  stg.file_of = file_of;
  stg.coef_bytes = b->coef_total * sizeof(int16_t);
  // ---- device arenas ----
  size_t off = 0;
  auto take = [&](size_t bytes) {
    const size_t o = off;
    off += Align(bytes, 256);
    return o;
  };
  const size_t o_stream = take(stage_bytes);
  const size_t o_state = take(n_sub_total * 8), o_used = take(n_sub_total * 8), o_nblk = take(n_sub_total * 4),
               o_loc = take(n_sub_total * 4);
  const size_t o_bsum = take(nb * 4), o_bflag = take(nb * 4), o_carry = take(nb * 4), o_bds = take(nb * sizeof(HuffBlockDesc));
  const size_t o_fds = take(nf * sizeof(HuffFileDesc)), o_tabs = take(tabs.size() * sizeof(DevHuffTable));
  const size_t o_dc = take(dc_total * 2), o_err = take(nf * 4 + 8), o_dcp = take(static_cast<size_t>(nf) * 3 * 8 * 2 * 4);
  {
    auto grow = [&](void** p, size_t* cap, size_t need) -> cudaError_t {
      if (need <= *cap) return cudaSuccess;
      cudaError_t e = cudaStreamSynchronize(compute_);
      if (e != cudaSuccess) return e;
      if (*p) cudaFree(*p);
      *p = nullptr;
      *cap = 0;
      e = cudaMalloc(p, need + need / 4);
      if (e == cudaSuccess) *cap = need + need / 4;
      return e;
    };
    // (the compute stream is idle whenever an upload is enqueued: the previous batch has been synchronised and this
    // batch's kernels come later, so growing an arena here never pulls memory from under a running kernel)
    RN_CUDA(grow(reinterpret_cast<void**>(&d_huff_[slot]), &d_huff_cap_[slot], off));
    RN_CUDA(GrowJpegBuffers(b->coef_total * sizeof(int16_t), 0, 0, 0, 0));
    if (!h_huff_flags_) RN_CUDA(cudaMallocHost(reinterpret_cast<void**>(&h_huff_flags_), 64 + max_batch_ * sizeof(int)));
    for (auto& ev : ev_huff_up_)
      if (!ev) RN_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  }
  // The upload runs on the copy stream: while the kernels of the previous batch work on the compute stream, this
  // batch's compressed bytes cross PCIe into the other arena.
  uint8_t* base = d_huff_[slot];
  RN_CUDA(cudaMemcpyAsync(base + o_stream, h_stream, stage_bytes, cudaMemcpyHostToDevice, copy_));
  RN_CUDA(cudaMemcpyAsync(base + o_bds, bds.data(), nb * sizeof(HuffBlockDesc), cudaMemcpyHostToDevice, copy_));
  RN_CUDA(cudaMemcpyAsync(base + o_fds, fds.data(), nf * sizeof(HuffFileDesc), cudaMemcpyHostToDevice, copy_));
  RN_CUDA(cudaMemcpyAsync(base + o_tabs, tabs.data(), tabs.size() * sizeof(DevHuffTable), cudaMemcpyHostToDevice, copy_));
  RN_CUDA(cudaMemsetAsync(base + o_err, 0, nf * 4 + 8, copy_));
  RN_CUDA(cudaEventRecord(ev_huff_up_[slot], copy_));
  stg.o_err = o_err;
  HuffBatch& hb = stg.hb;
  hb = HuffBatch{};
  hb.files = reinterpret_cast<const HuffFileDesc*>(base + o_fds);
  hb.blocks = reinterpret_cast<const HuffBlockDesc*>(base + o_bds);
  hb.tables = reinterpret_cast<const DevHuffTable*>(base + o_tabs);
  hb.streams = base + o_stream;
  hb.sub_seg = reinterpret_cast<const int*>(base + o_stream + stream_total);
  hb.state = reinterpret_cast<unsigned long long*>(base + o_state);
  hb.start_used = reinterpret_cast<unsigned long long*>(base + o_used);
  hb.nblk = reinterpret_cast<unsigned*>(base + o_nblk);
  hb.local_off = reinterpret_cast<unsigned*>(base + o_loc);
  hb.block_sum = reinterpret_cast<unsigned*>(base + o_bsum);
  hb.block_has_start = reinterpret_cast<int*>(base + o_bflag);
  hb.carry = reinterpret_cast<unsigned*>(base + o_carry);
  hb.coefs = d_coef_;
  hb.dcdiff = reinterpret_cast<int16_t*>(base + o_dc);
  hb.dc_part = reinterpret_cast<int*>(base + o_dcp);
  hb.file_error = reinterpret_cast<int*>(base + o_err);
  hb.changed = reinterpret_cast<int*>(base + o_err) + nf;
  hb.h_changed = h_huff_flags_;
  hb.n_files = nf;
  hb.n_blocks = nb;
  hb.rounds_out = &jpeg_huffman_rounds_;
  return cudaSuccess;
}

// Second half of the stage: the Huffman kernels of the batch uploaded into `slot` (enqueued on compute_).
cudaError_t Replica::JpegHuffRun(JpegBatch* b, int slot, std::vector<int>* err) {
  const int m = static_cast<int>(b->index.size());
  err->assign(m, 0);
  HuffStage& stg = huff_stage_[slot];
  huff_file_of_.clear();
  if (!stg.any) return cudaSuccess;
  RN_CUDA(cudaStreamWaitEvent(compute_, ev_huff_up_[slot], 0));
  RN_CUDA(cudaMemsetAsync(d_coef_, 0, stg.coef_bytes, compute_));
  {
    cudaError_t e = HuffDecode(stg.hb, compute_);
    if (e != cudaSuccess) {
      err_ = std::string("device Huffman decode: ") + cudaGetErrorString(e);
      return e;
    }
  }
  if (jpeg_huffman_rounds_ < 0) {  // no fixed point within the launch budget: every file of the batch goes to the host path
    for (int k = 0; k < m; ++k)
      if (b->st[k] == kJpegOk) b->st[k] = kJpegUnsupported;
    return cudaSuccess;
  }
  last_launches_ += 4 + jpeg_huffman_rounds_ + 1;
  // error flags travel back with the rest of the batch (valid after the caller's stream synchronisation)
  RN_CUDA(cudaMemcpyAsync(h_huff_flags_ + 16, d_huff_[slot] + stg.o_err, stg.nf * sizeof(int), cudaMemcpyDeviceToHost, compute_));
  huff_file_of_ = stg.file_of;
  return cudaSuccess;
}

// after the stream has been synchronised: which batch entries did the device report as damaged
void Replica::JpegHuffmanErrors(std::vector<int>* err) const {
  for (size_t j = 0; j < huff_file_of_.size(); ++j)
    if (h_huff_flags_[16 + j]) (*err)[huff_file_of_[j]] = 1;
}

cudaError_t Replica::DecodeJpeg(const uint8_t* file, size_t size, uint8_t* out, size_t capacity, int* height, int* width,
                                int32_t* status) {
  {
    cudaError_t ew = WaitHost(~0ull);
    if (ew != cudaSuccess) return ew;
  }
  RN_CUDA(cudaSetDevice(device_));
  JpegBatch b;
  b.info.resize(1);
  b.index.assign(1, 0);
  const JpegStatus hs = JpegParseHeader(file, size, &b.info[0]);
  *status = hs;
  if (hs != kJpegOk) return cudaSuccess;
  const JpegInfo& f = b.info[0];
  const bool swap = f.orientation >= 5;
  *height = swap ? f.width : f.height;
  *width = swap ? f.height : f.width;
  const size_t bytes = static_cast<size_t>(f.width) * f.height * 3;
  if (!out) return cudaSuccess;
  if (capacity < bytes) {
    err_ = "output buffer smaller than height * width * 3";
    return cudaErrorInvalidValue;
  }
  b.coef_off.assign(1, 0);
  b.coef_total = Align(f.coef_count, 64);
  last_launches_ = 0;
  std::vector<CropDesc> crops;
  std::vector<char> ok;
  const size_t sz = size;
  cudaError_t e = cudaSuccess;
  bool on_device = false;
  if (!(flags_ & RN_FLAG_JPEG_HOST_HUFFMAN)) {
    std::vector<int> err;
    e = JpegHuffUpload(&file, &sz, &b, 1, nullptr, 0);
    if (e != cudaSuccess) return e;
    e = JpegHuffRun(&b, 0, &err);
    if (e != cudaSuccess) return e;
    if (b.st[0] == kJpegOk) {
      e = JpegToRaw(b, nullptr, &crops, &ok, status);
      if (e != cudaSuccess) return e;
      RN_CUDA(cudaStreamSynchronize(compute_));
      JpegHuffmanErrors(&err);
      on_device = err[0] == 0;
      if (on_device) ++jpeg_device_huffman_files_;
    }
  }
  if (!on_device) {
    e = GrowJpegBuffers(b.coef_total * sizeof(int16_t), 0, 0, 1, 1);
    if (e != cudaSuccess) return e;
    RN_CUDA(cudaStreamSynchronize(compute_));  // an earlier upload may still be reading the pinned buffer
    JpegDecodeHost(&file, &sz, &b, h_coef_[0], 1);
    e = JpegToRaw(b, h_coef_[0], &crops, &ok, status);
    if (e != cudaSuccess) return e;
    ++jpeg_host_huffman_files_;
  }
  if (!ok[0]) return cudaSuccess;
  RN_CUDA(cudaMemcpy2DAsync(out, static_cast<size_t>(*width) * 3, d_raw_ + crops[0].offset,
                            static_cast<size_t>(crops[0].W) * 3, static_cast<size_t>(*width) * 3, *height,
                            cudaMemcpyDeviceToHost, compute_));
  RN_CUDA(cudaStreamSynchronize(compute_));
  return cudaSuccess;
}

cudaError_t Replica::InferJpegsHostHuffman(const uint8_t* const* files, const size_t* sizes, int n, int threads,
                                           int64_t* top1, float* probs, float* logits, int32_t* status) {
  NvtxRangeJpeg nvtx_range("rn::InferJpegs (entropy decode on host threads | IDCT + colour + crop + forward on the device)");
  const int S = shape_.im_side, C = shape_.num_classes;
  if (!d_descs_) RN_CUDA(Alloc(&d_descs_, static_cast<size_t>(max_batch_) * sizeof(CropDesc)));
  // headers first: geometry decides the micro-batches (at most max_batch files and ~0.4 GB of device staging each)
  std::vector<JpegBatch> batches;
  {
    constexpr size_t kStagingCap = size_t{3} << 27;  // 384 MB: a handful of photographs, so the pipeline fills quickly
    JpegInfo f;
    size_t staged = 0;
    for (int i = 0; i < n; ++i) {
      status[i] = JpegParseHeader(files[i], sizes[i], &f);
      if (status[i] != kJpegOk) continue;
      const size_t need = f.coef_count * 3 + static_cast<size_t>(f.width) * f.height * 3 + 4096;
      if (batches.empty() || static_cast<int>(batches.back().index.size()) >= max_batch_ ||
          staged + need > kStagingCap) {
        batches.emplace_back();
        staged = 0;
      }
      staged += need;
      JpegBatch& b = batches.back();
      b.index.push_back(i);
      b.info.push_back(f);
      b.coef_off.push_back(b.coef_total);
      b.coef_total += Align(f.coef_count, 64);
    }
  }
  if (batches.empty()) return cudaSuccess;
  size_t max_coef = 0;
  for (const auto& b : batches) max_coef = std::max(max_coef, b.coef_total);
  const int nb = static_cast<int>(batches.size());
  {
    cudaError_t e = GrowJpegBuffers(max_coef * sizeof(int16_t), 0, 0, 0, std::min(nb, kJpegRing));
    if (e != cudaSuccess) return e;
  }
  RN_CUDA(cudaStreamSynchronize(compute_));  // an earlier upload may still be reading the pinned buffers

  // Host threads walk the files of ALL batches in order (so small batches do not idle them) and run up to
  // kJpegRing - 1 batches ahead of the device: batch b is entropy-decoded into pinned buffer b % kJpegRing, which is
  // free once the device work of batch b - kJpegRing has completed.
  std::mutex mu;
  std::condition_variable cv;
  int released = 0;  // batches whose pinned buffer may be overwritten
  bool abort = false;
  std::unique_ptr<std::atomic<int>[]> done(new std::atomic<int>[nb]);
  std::vector<std::pair<int, int>> items;
  for (int b = 0; b < nb; ++b) {
    done[b].store(0);
    batches[b].st.assign(batches[b].index.size(), kJpegOk);
    for (int k = 0; k < static_cast<int>(batches[b].index.size()); ++k) items.emplace_back(b, k);
  }
  std::atomic<int> next{0};
  auto work = [&]() {
    for (int i = next.fetch_add(1); i < static_cast<int>(items.size()); i = next.fetch_add(1)) {
      const int b = items[i].first, k = items[i].second;
      {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&] { return abort || released >= b - (kJpegRing - 1); });
        if (abort) return;
      }
      JpegBatch& bt = batches[b];
      bt.st[k] = JpegDecodeCoefficients(files[bt.index[k]], sizes[bt.index[k]], bt.info[k], h_coef_[b % kJpegRing] + bt.coef_off[k]);
      if (done[b].fetch_add(1) + 1 == static_cast<int>(bt.index.size())) {
        { std::lock_guard<std::mutex> lk(mu); }
        cv.notify_all();
      }
    }
  };
  std::vector<std::thread> pool;
  const int nt = std::max(1, std::min(threads, static_cast<int>(items.size())));
  for (int t = 0; t < nt; ++t) pool.emplace_back(work);
  auto stop = [&](cudaError_t e) {
    {
      std::lock_guard<std::mutex> lk(mu);
      abort = true;
    }
    cv.notify_all();
    for (auto& t : pool) t.join();
    return e;
  };
  std::vector<CropDesc> crops, packed;
  std::vector<char> ok;
  std::vector<int> where;
  for (int b = 0; b < nb; ++b) {
    {
      std::unique_lock<std::mutex> lk(mu);
      cv.wait(lk, [&] { return done[b].load() == static_cast<int>(batches[b].index.size()); });
    }
    cudaError_t e = JpegToRaw(batches[b], h_coef_[b % kJpegRing], &crops, &ok, status);
    if (e != cudaSuccess) return stop(e);
    packed.clear();
    where.clear();
    for (size_t k = 0; k < batches[b].index.size(); ++k) {
      if (!ok[k]) continue;
      packed.push_back(crops[k]);
      where.push_back(batches[b].index[k]);
    }
    const int m = static_cast<int>(packed.size());
    const HostOut ho = Out(0);
    if (m > 0) {
      if ((e = cudaMemcpyAsync(d_descs_, packed.data(), m * sizeof(CropDesc), cudaMemcpyHostToDevice, compute_)) != cudaSuccess ||
          (e = CropResizeBatchU8(d_raw_, static_cast<const CropDesc*>(d_descs_), m, static_cast<uint8_t*>(d_in_[0]), S,
                                 compute_)) != cudaSuccess) {
        err_ = std::string("InferJpegs: ") + cudaGetErrorString(e);
        return stop(e);
      }
      ++last_launches_;
      cur_ = &sets_[0];
      e = ForwardDevice(d_in_[0], InputKind::kU8Bgr, m, ho.top1, ho.probs, ho.logits, compute_);
      if (e != cudaSuccess) return stop(e);
    }
    if ((e = cudaStreamSynchronize(compute_)) != cudaSuccess) {
      err_ = std::string("InferJpegs: ") + cudaGetErrorString(e);
      return stop(e);
    }
    {
      std::lock_guard<std::mutex> lk(mu);
      released = b + 1;
    }
    cv.notify_all();
    for (int k = 0; k < m; ++k) {  // scatter: files that fell out keep their slots untouched
      const int dst = where[k];
      if (top1) top1[dst] = ho.top1[k];
      if (probs) std::memcpy(probs + static_cast<size_t>(dst) * C, ho.probs + static_cast<size_t>(k) * C, C * sizeof(float));
      if (logits) std::memcpy(logits + static_cast<size_t>(dst) * C, ho.logits + static_cast<size_t>(k) * C, C * sizeof(float));
    }
  }
  for (auto& t : pool) t.join();
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
  std::vector<int> todo;  // files for the host Huffman decoder
  if (flags_ & RN_FLAG_JPEG_HOST_HUFFMAN) {
    for (int i = 0; i < n; ++i) todo.push_back(i);
  } else {
    NvtxRangeJpeg nvtx_range("rn::InferJpegs (Huffman decode + IDCT + colour + crop + forward on the device)");
    const int S = shape_.im_side, C = shape_.num_classes;
    if (!d_descs_) RN_CUDA(Alloc(&d_descs_, static_cast<size_t>(max_batch_) * sizeof(CropDesc)));
    std::vector<JpegBatch> batches;
    {
      constexpr size_t kStagingCap = size_t{3} << 29;  // 1.5 GB of device staging per batch
      JpegInfo f;
      size_t staged = 0;
      for (int i = 0; i < n; ++i) {
        status[i] = JpegParseHeader(files[i], sizes[i], &f);
        if (status[i] != kJpegOk) continue;
        const size_t need = f.coef_count * 3 + static_cast<size_t>(f.width) * f.height * 3 + sizes[i] * 2 + 4096;
        if (batches.empty() || static_cast<int>(batches.back().index.size()) >= max_batch_ || staged + need > kStagingCap) {
          batches.emplace_back();
          staged = 0;
        }
        staged += need;
        JpegBatch& b = batches.back();
        b.index.push_back(i);
        b.info.push_back(f);
        b.coef_off.push_back(b.coef_total);
        b.coef_total += Align(f.coef_count, 64);
      }
    }
    // Host threads strip the byte stuffing of the files of ALL batches in order, into a ring of pinned staging buffers,
    // up to kJpegRing - 1 batches ahead of the device.
    const int nbat = static_cast<int>(batches.size());
    size_t max_stage = 0;
    for (auto& b : batches) {
      b.LayOutStreams(sizes);
      max_stage = std::max(max_stage, b.stage_bytes);
    }
    if (nbat > 0) {
      cudaError_t e = GrowJpegBuffers((max_stage + 1) / 2 * 2, 0, 0, 0, std::min(nbat, kJpegRing));
      if (e != cudaSuccess) return e;
      RN_CUDA(cudaStreamSynchronize(compute_));
      RN_CUDA(cudaStreamSynchronize(copy_));
    }
    std::mutex mu;
    std::condition_variable cv;
    int released = 0;
    bool abort = false;
    std::unique_ptr<std::atomic<int>[]> done(new std::atomic<int>[std::max(nbat, 1)]);
    std::vector<std::pair<int, int>> items;
    for (int bi = 0; bi < nbat; ++bi) {
      done[bi].store(0);
      for (int k = 0; k < static_cast<int>(batches[bi].index.size()); ++k) items.emplace_back(bi, k);
    }
    std::atomic<int> next{0};
    auto work = [&]() {
      for (int i = next.fetch_add(1); i < static_cast<int>(items.size()); i = next.fetch_add(1)) {
        const int bi = items[i].first, k = items[i].second;
        {
          std::unique_lock<std::mutex> lk(mu);
          cv.wait(lk, [&] { return abort || released >= bi - (kJpegRing - 1); });
          if (abort) return;
        }
        batches[bi].PrepareStream(files, sizes, k, reinterpret_cast<uint8_t*>(h_coef_[bi % kJpegRing]));
        if (done[bi].fetch_add(1) + 1 == static_cast<int>(batches[bi].index.size())) {
          { std::lock_guard<std::mutex> lk(mu); }
          cv.notify_all();
        }
      }
    };
    std::vector<std::thread> pool;
    for (int t = 0; t < std::max(1, std::min(threads, static_cast<int>(items.size()))) && !items.empty(); ++t)
      pool.emplace_back(work);
    struct Joiner {  // every exit path stops and joins the pool
      std::mutex& mu;
      std::condition_variable& cv;
      bool& abort;
      std::vector<std::thread>& pool;
      ~Joiner() {
        {
          std::lock_guard<std::mutex> lk(mu);
          abort = true;
        }
        cv.notify_all();
        for (auto& t : pool) t.join();
      }
    } joiner{mu, cv, abort, pool};
    std::vector<CropDesc> crops, packed;
    std::vector<char> ok;
    std::vector<int> where, entry, err;
    for (int bi = 0; bi < nbat; ++bi) {
      JpegBatch& b = batches[bi];
      {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&] { return done[bi].load() == static_cast<int>(b.index.size()); });
      }
      struct Release {  // hand the staging buffer back when this batch is done, whatever happens
        std::mutex& mu;
        std::condition_variable& cv;
        int& released;
        int value;
        ~Release() {
          {
            std::lock_guard<std::mutex> lk(mu);
            released = value;
          }
          cv.notify_all();
        }
      } release{mu, cv, released, bi + 1};
      // this batch was uploaded during the previous iteration; upload the next one now, so that its bytes cross PCIe
      // while this batch's kernels run
      cudaError_t e = cudaSuccess;
      if (bi == 0) {
        e = JpegHuffUpload(files, sizes, &b, threads, reinterpret_cast<const uint8_t*>(h_coef_[0]), 0);
        if (e != cudaSuccess) return e;
      }
      if (bi + 1 < nbat) {
        JpegBatch& nx = batches[bi + 1];
        {
          std::unique_lock<std::mutex> lk(mu);
          cv.wait(lk, [&] { return done[bi + 1].load() == static_cast<int>(nx.index.size()); });
        }
        e = JpegHuffUpload(files, sizes, &nx, threads, reinterpret_cast<const uint8_t*>(h_coef_[(bi + 1) % kJpegRing]),
                           (bi + 1) & 1);
        if (e != cudaSuccess) return e;
      }
      e = JpegHuffRun(&b, bi & 1, &err);
      if (e != cudaSuccess) return e;
      // entries the device did not take: JpegToRaw must skip them, the host path picks them up below
      e = JpegToRaw(b, nullptr, &crops, &ok, nullptr);
      if (e != cudaSuccess) return e;
      packed.clear();
      where.clear();
      entry.clear();
      for (size_t k = 0; k < b.index.size(); ++k) {
        if (!ok[k]) {
          todo.push_back(b.index[k]);
          continue;
        }
        packed.push_back(crops[k]);
        where.push_back(b.index[k]);
        entry.push_back(static_cast<int>(k));
      }
      const int m = static_cast<int>(packed.size());
      if (m == 0) continue;
      const HostOut ho = Out(0);
      RN_CUDA(cudaMemcpyAsync(d_descs_, packed.data(), m * sizeof(CropDesc), cudaMemcpyHostToDevice, compute_));
      RN_CUDA(CropResizeBatchU8(d_raw_, static_cast<const CropDesc*>(d_descs_), m, static_cast<uint8_t*>(d_in_[0]), S, compute_));
      ++last_launches_;
      cur_ = &sets_[0];
      e = ForwardDevice(d_in_[0], InputKind::kU8Bgr, m, ho.top1, ho.probs, ho.logits, compute_);
      if (e != cudaSuccess) return e;
      RN_CUDA(cudaStreamSynchronize(compute_));
      JpegHuffmanErrors(&err);
      for (int k = 0; k < m; ++k) {
        const int dst = where[k];
        if (err[entry[k]]) {  // the device found the stream damaged: the host decoder has the last word on this file
          todo.push_back(dst);
          continue;
        }
        ++jpeg_device_huffman_files_;
        status[dst] = kJpegOk;
        if (top1) top1[dst] = ho.top1[k];
        if (probs) std::memcpy(probs + static_cast<size_t>(dst) * C, ho.probs + static_cast<size_t>(k) * C, C * sizeof(float));
        if (logits) std::memcpy(logits + static_cast<size_t>(dst) * C, ho.logits + static_cast<size_t>(k) * C, C * sizeof(float));
      }
    }
  }
  if (todo.empty()) return cudaSuccess;
  // ---- the rest: Huffman decoding on host threads ----
  std::sort(todo.begin(), todo.end());
  const int C = shape_.num_classes;
  const int r = static_cast<int>(todo.size());
  std::vector<const uint8_t*> f2(r);
  std::vector<size_t> s2(r);
  std::vector<int64_t> t2(r, -1);
  std::vector<float> p2(static_cast<size_t>(r) * C), l2(static_cast<size_t>(r) * C);
  std::vector<int32_t> st2(r, kJpegUnsupported);
  for (int j = 0; j < r; ++j) {
    f2[j] = files[todo[j]];
    s2[j] = sizes[todo[j]];
  }
  cudaError_t e = InferJpegsHostHuffman(f2.data(), s2.data(), r, threads, t2.data(), p2.data(), l2.data(), st2.data());
  if (e != cudaSuccess) return e;
  for (int j = 0; j < r; ++j) {
    const int dst = todo[j];
    status[dst] = st2[j];
    if (st2[j] != kJpegOk) continue;
    ++jpeg_host_huffman_files_;
    if (top1) top1[dst] = t2[j];
    if (probs) std::memcpy(probs + static_cast<size_t>(dst) * C, p2.data() + static_cast<size_t>(j) * C, C * sizeof(float));
    if (logits) std::memcpy(logits + static_cast<size_t>(dst) * C, l2.data() + static_cast<size_t>(j) * C, C * sizeof(float));
  }
  return cudaSuccess;
}

}  // namespace rn
