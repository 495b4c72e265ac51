// TensorFlow "V2 bundle" checkpoint reader (product code, host side).
// Replaces Saver.restore (reference network.py:122) for the inference path.
#pragma once
#include <cstdint>
#include <map>
#include <string>
#include <vector>

namespace rn {

struct Tensor {
  std::vector<int64_t> shape;
  std::vector<float> data;
  int64_t numel() const {
    int64_t n = 1;
    for (auto d : shape) n *= d;
    return n;
  }
};

using TensorMap = std::map<std::string, Tensor>;

enum class BundleError { kOk = 0, kIo, kFormat };

// Reads every DT_FLOAT tensor of <prefix>.index / <prefix>.data-00000-of-00001.
// Verifies the SSTable block CRCs and the per-tensor masked CRC32C.
BundleError ReadBundle(const std::string& prefix, TensorMap* out, std::string* err);

uint32_t Crc32c(const uint8_t* data, size_t n);
uint32_t MaskCrc(uint32_t crc);

}  // namespace rn
