// Network description + exact forward BatchNorm fold (host side, fp64).
//
// The reference applies conv -> ReLU6 -> AvgPool -> BN (network.py:184-194), so BN
// cannot be folded backwards through the ReLU6; because every conv is VALID
// (network.py:172) it folds exactly FORWARDS into the next conv/dense layer.
// Residual joins (network.py:199-203) become  A*p_k + B*resize(p_0) + C.
#pragma once
#include <string>
#include <vector>

#include "tf_bundle.h"

namespace rn {

constexpr int kNumConvs = 10;
constexpr int kNumDense = 4;

struct ConvShape {
  int cin, cout;
  int in_side;    // input spatial size
  int conv_side;  // after 3x3 VALID conv
  int out_side;   // after pool (== conv_side when pool_k == 0)
  int pool_k, pool_s;
  int join_src;  // index of the conv whose pooled output is the residual source, or -1
};

struct NetShape {
  int im_side = 0, num_classes = 0, flat_len = 0;
  ConvShape conv[kNumConvs];
  int dense_in[kNumDense], dense_out[kNumDense];
};

// Static structure of the graph (reference network.py:225-237).
bool MakeNetShape(int im_side, int num_classes, NetShape* out, std::string* err);

struct FoldedConv {
  std::vector<double> w;  // HWIO [3][3][cin][cout], BN of the producer folded in
  std::vector<double> b;  // [cout]
};
struct FoldedJoin {
  std::vector<double> a, b, c;  // per channel
};
struct FoldedDense {
  std::vector<double> w;  // [in][out]
  std::vector<double> b;  // [out]
};

struct FoldedNet {
  NetShape shape;
  // conv[0] comes in two flavours: fed with uint8 BGR pixels (normalisation
  // (x/255)*2-1 and the BGR->RGB swap folded in, network.py:129/153) or with the raw
  // float RGB feed of sess.run (network.py:155).
  FoldedConv conv0_u8bgr, conv0_u8rgb, conv0_f32rgb;
  FoldedConv conv[kNumConvs];  // conv[0] unused (see above)
  FoldedJoin join[kNumConvs];  // valid where shape.conv[i].join_src >= 0
  FoldedDense dense[kNumDense];
};

// `dense0` (optional) overrides dense/kernel for im_side != 224.
bool FoldNetwork(const TensorMap& vars, const NetShape& shape, const Tensor* dense0, FoldedNet* out,
                 std::string* err);

}  // namespace rn
