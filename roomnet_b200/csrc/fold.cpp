#include "fold.h"

#include <cmath>

namespace rn {
namespace {

// reference network.py:226-230: (filters, pooling, pool_ksize, pool_stride, block_depth)
struct BlockSpec {
  int filters, pool_k, pool_s, depth;
};
const BlockSpec kBlocks[5] = {{8, 3, 1, 1}, {32, 4, 1, 3}, {64, 4, 2, 2}, {128, 0, 0, 1}, {16, 4, 2, 3}};
const int kDenseUnits[3] = {32, 16, 8};  // reference network.py:234-236
// attr "epsilon" of every FusedBatchNorm node / batchnorm/add/y const in final_model/roomnet.meta
const double kBnEps = static_cast<double>(0.0010000000474974513f);

std::string Suffixed(const char* base, int i) { return i == 0 ? std::string(base) : std::string(base) + "_" + std::to_string(i); }

const Tensor* Find(const TensorMap& m, const std::string& name, std::vector<int64_t> shape, std::string* err) {
  auto it = m.find(name);
  if (it == m.end()) {
    *err = "checkpoint has no tensor '" + name + "'";
    return nullptr;
  }
  if (it->second.shape != shape) {
    std::string got, want;
    for (auto d : it->second.shape) got += std::to_string(d) + ",";
    for (auto d : shape) want += std::to_string(d) + ",";
    *err = "tensor '" + name + "' has shape [" + got + "] but the graph needs [" + want + "]";
    return nullptr;
  }
  return &it->second;
}

struct Affine {
  std::vector<double> s, t;
  bool valid = false;
};

bool BnAffine(const TensorMap& m, int idx, int ch, Affine* out, std::string* err) {
  std::string n = Suffixed("batch_normalization", idx);
  const Tensor* g = Find(m, n + "/gamma", {ch}, err);
  const Tensor* b = g ? Find(m, n + "/beta", {ch}, err) : nullptr;
  const Tensor* mu = b ? Find(m, n + "/moving_mean", {ch}, err) : nullptr;
  const Tensor* var = mu ? Find(m, n + "/moving_variance", {ch}, err) : nullptr;
  if (!var) return false;
  out->s.resize(ch);
  out->t.resize(ch);
  for (int c = 0; c < ch; ++c) {
    double s = static_cast<double>(g->data[c]) / std::sqrt(static_cast<double>(var->data[c]) + kBnEps);
    out->s[c] = s;
    out->t[c] = static_cast<double>(b->data[c]) - static_cast<double>(mu->data[c]) * s;
  }
  out->valid = true;
  return true;
}

}  // namespace

bool MakeNetShape(int im_side, int num_classes, NetShape* out, std::string* err) {
  if (num_classes < 1 || num_classes > 64) {
    *err = "num_classes out of range";
    return false;
  }
  out->im_side = im_side;
  out->num_classes = num_classes;
  int s = im_side, cin = 3, ci = 0;
  for (const auto& blk : kBlocks) {
    int first = ci;
    for (int d = 0; d < blk.depth; ++d, ++ci) {
      ConvShape& c = out->conv[ci];
      c.cin = cin;
      c.cout = blk.filters;
      c.in_side = s;
      c.conv_side = s - 2;
      c.pool_k = blk.pool_k;
      c.pool_s = blk.pool_s;
      c.out_side = blk.pool_k ? (c.conv_side - blk.pool_k) / blk.pool_s + 1 : c.conv_side;
      c.join_src = -1;
      if (c.conv_side < 1 || (blk.pool_k && c.conv_side < blk.pool_k)) {
        *err = "im_side " + std::to_string(im_side) + " is too small for the RoomNet layer stack";
        return false;
      }
      s = c.out_side;
      cin = blk.filters;
    }
    if (blk.depth > 1) out->conv[ci - 1].join_src = first;
  }
  out->flat_len = s * s * cin;
  int in = out->flat_len;
  for (int i = 0; i < kNumDense; ++i) {
    out->dense_in[i] = in;
    out->dense_out[i] = i < 3 ? kDenseUnits[i] : num_classes;
    in = out->dense_out[i];
  }
  return true;
}

bool FoldNetwork(const TensorMap& vars, const NetShape& shape, const Tensor* dense0, FoldedNet* out,
                 std::string* err) {
  out->shape = shape;
  int bn = 0;
  Affine pending;  // affine of the BN that feeds the next conv/dense (valid=false: input is a join output)
  Affine post[kNumConvs];  // affine of the BN that follows conv i's pool
  for (int i = 0; i < kNumConvs; ++i) {
    const ConvShape& cs = shape.conv[i];
    const Tensor* k = Find(vars, Suffixed("conv2d", i) + "/kernel", {3, 3, cs.cin, cs.cout}, err);
    if (!k) return false;
    auto W = [&](int tap, int c, int o) { return static_cast<double>(k->data[(tap * cs.cin + c) * cs.cout + o]); };
    if (i == 0) {
      // x_rgb[c] = v[c']*(2/255) - 1 with c' = 2-c for BGR bytes, c for RGB bytes (network.py:129/153).
      FoldedConv &bgr = out->conv0_u8bgr, &rgb = out->conv0_u8rgb, &f32 = out->conv0_f32rgb;
      for (FoldedConv* f : {&bgr, &rgb, &f32}) {
        f->w.assign(9 * 3 * cs.cout, 0.0);
        f->b.assign(cs.cout, 0.0);
      }
      for (int tap = 0; tap < 9; ++tap)
        for (int c = 0; c < 3; ++c)
          for (int o = 0; o < cs.cout; ++o) {
            double w = W(tap, c, o);
            f32.w[(tap * 3 + c) * cs.cout + o] = w;
            rgb.w[(tap * 3 + c) * cs.cout + o] = w * (2.0 / 255.0);
            bgr.w[(tap * 3 + (2 - c)) * cs.cout + o] = w * (2.0 / 255.0);
            rgb.b[o] -= w;
            bgr.b[o] -= w;
          }
    } else {
      FoldedConv& f = out->conv[i];
      f.w.resize(9 * cs.cin * cs.cout);
      f.b.assign(cs.cout, 0.0);
      for (int tap = 0; tap < 9; ++tap)
        for (int c = 0; c < cs.cin; ++c)
          for (int o = 0; o < cs.cout; ++o) {
            double w = W(tap, c, o);
            if (pending.valid) {
              f.b[o] += w * pending.t[c];
              w *= pending.s[c];
            }
            f.w[(tap * cs.cin + c) * cs.cout + o] = w;
          }
    }
    if (!BnAffine(vars, bn++, cs.cout, &pending, err)) return false;
    post[i] = pending;
    if (cs.join_src >= 0) {
      const Affine &first_affine = post[cs.join_src], &last_affine = post[i];
      Affine r;
      if (!BnAffine(vars, bn++, cs.cout, &r, err)) return false;
      FoldedJoin& j = out->join[i];
      j.a.resize(cs.cout);
      j.b.resize(cs.cout);
      j.c.resize(cs.cout);
      for (int c = 0; c < cs.cout; ++c) {
        j.a[c] = r.s[c] * last_affine.s[c];
        j.b[c] = r.s[c] * first_affine.s[c];
        j.c[c] = r.s[c] * (last_affine.t[c] + first_affine.t[c]) + r.t[c];
      }
      pending.valid = false;
    }
  }
  for (int i = 0; i < kNumDense; ++i) {
    int in = shape.dense_in[i], on = shape.dense_out[i];
    const Tensor* k = (i == 0 && dense0) ? dense0 : Find(vars, Suffixed("dense", i) + "/kernel", {in, on}, err);
    if (!k) return false;
    if (k->shape != std::vector<int64_t>{in, on}) {
      *err = "dense/kernel override has the wrong shape for im_side " + std::to_string(shape.im_side);
      return false;
    }
    FoldedDense& f = out->dense[i];
    f.w.resize(static_cast<size_t>(in) * on);
    f.b.assign(on, 0.0);
    for (int r = 0; r < in; ++r)
      for (int o = 0; o < on; ++o) {
        double w = static_cast<double>(k->data[static_cast<size_t>(r) * on + o]);
        if (pending.valid) {
          f.b[o] += w * pending.t[r];
          w *= pending.s[r];
        }
        f.w[static_cast<size_t>(r) * on + o] = w;
      }
    if (i == kNumDense - 1) {  // dense_3 is the only biased layer (network.py:237)
      const Tensor* bias = Find(vars, Suffixed("dense", i) + "/bias", {on}, err);
      if (!bias) return false;
      for (int o = 0; o < on; ++o) f.b[o] += static_cast<double>(bias->data[o]);
    } else {
      if (!BnAffine(vars, bn++, on, &pending, err)) return false;
    }
  }
  return true;
}

}  // namespace rn
