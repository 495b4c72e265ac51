// Host half of the JPEG front end (replaces cv2.imread of the reference's classify_im_dir, infer.py:81, for baseline
// JPEG files).  Default: the host parses the markers, puts the Huffman tables into device form and strips the byte
// stuffing of the scan (JpegPrepareScan); Huffman decoding (kernels_jpeg_huff.cu), dequantisation, inverse DCT, chroma
// upsampling and colour conversion (kernels_jpeg.cu) are CUDA kernels with the integer arithmetic of libjpeg-turbo's
// default decoder (JDCT_ISLOW, fancy upsampling), i.e. bit-identical to cv2.imread.  Fallback (files with several
// scans, RN_FLAG_JPEG_HOST_HUFFMAN) and test oracle: a complete table-driven Huffman decoder on the host
// (JpegDecodeCoefficients), whose quantised coefficients go to the same device kernels.
//
// Supported: 8-bit baseline / extended-sequential Huffman JPEG (SOF0 / SOF1), grey or YCbCr, luma sampling 1x1, 2x1,
// 2x2 with 1x1 chroma, any number of scans, restart intervals, EXIF orientation.  Everything else (progressive,
// arithmetic coding, CMYK, 12-bit, exotic sampling, damaged streams) is reported as kJpegUnsupported so that the caller
// decodes that file on the host instead.
#pragma once
#include <cstddef>
#include <cstdint>

namespace rn {

enum JpegStatus : int { kJpegOk = 0, kJpegUnsupported = 1, kJpegCorrupt = 2 };

struct JpegComponent {
  int id = 0;
  int h = 1, v = 1;    // sampling factors
  int tq = 0;          // quantisation table
  int wblocks = 0;     // ceil(downsampled width / 8): blocks that carry image data
  int hblocks = 0;
  int dw = 0, dh = 0;  // downsampled width / height in samples (libjpeg: downsampled_width / _height)
  size_t coef_offset = 0;  // first coefficient of this component in the image's coefficient array (int16 units)
};

struct JpegInfo {
  int width = 0, height = 0;
  int ncomp = 0;
  int hmax = 1, vmax = 1;
  int orientation = 1;  // EXIF tag 0x0112 (1..8), 1 when absent
  int restart_interval = 0;
  JpegComponent comp[3];
  uint16_t quant[4][64];  // natural (row-major) order
  bool have_quant[4] = {false, false, false, false};
  size_t coef_count = 0;  // int16 coefficients of the whole image (sum over components of wblocks*hblocks*64)
  size_t sos_offset = 0;  // internal: where the first scan starts
};

// Header pass: fills `info` (geometry, tables, orientation); no entropy decoding.
JpegStatus JpegParseHeader(const uint8_t* data, size_t size, JpegInfo* info);
// Entropy-decodes every scan into `coefs` (info.coef_count int16 values, zero-initialised by this call): per component
// a [hblocks][wblocks][64] array in natural order, NOT dequantised.
JpegStatus JpegDecodeCoefficients(const uint8_t* data, size_t size, const JpegInfo& info, int16_t* coefs);

// ---- device Huffman decoding (kernels_jpeg_huff.cu): what the host prepares per file ----
// A Huffman table in the form the device decoder reads: 10-bit direct lookup, canonical-code search for longer codes.
struct DevHuffTable {
  uint16_t fast[1024];  // (length << 8) | symbol for codes of <= 10 bits, 0 = longer (or invalid) code
  int32_t maxcode[18];  // largest code of each length, -1 = none
  int32_t valoff[17];   // index of the first symbol of each length minus its first code
  uint8_t vals[256];
};
constexpr int kSubseqBytes = 128;  // the device decodes the stream in subsequences of 1024 bits, one thread each

// One file whose entropy-coded data is a single scan over all components (what encoders of photographs write).
struct JpegScanPlan {
  int bpm = 0;                 // blocks per MCU
  uint8_t blk_comp[8] = {};    // block j of an MCU: component, column and row inside the MCU
  uint8_t blk_hh[8] = {};
  uint8_t blk_vv[8] = {};
  int mcus_x = 0, mcus_y = 0;
  uint32_t total_blocks = 0;   // bpm * mcus_x * mcus_y, dummy edge blocks included
  uint32_t seg_blocks = 0;     // blocks per restart segment (= total_blocks when the file has no restart markers)
  int n_seg = 1;
  DevHuffTable tab[6];         // [2 * component] = DC table, [2 * component + 1] = AC table
  size_t stream_bytes = 0;     // unstuffed stream written by JpegPrepareScan: a multiple of kSubseqBytes
};
// Upper bound of JpegScanPlan::stream_bytes for a file of `size` bytes (known from the header alone).
size_t JpegStreamCapacity(const JpegInfo& info, size_t size);
// Removes the byte stuffing and the restart markers of the scan: every restart segment starts on a subsequence
// boundary of `stream` (padded with zero bytes), sub_seg[i] = restart segment of subsequence i.  kJpegOk = the device
// can decode this file; anything else = use JpegDecodeCoefficients (several scans, tables redefined, damaged stream ...).
JpegStatus JpegPrepareScan(const uint8_t* data, size_t size, const JpegInfo& info, JpegScanPlan* plan, uint8_t* stream,
                           size_t stream_capacity, int32_t* sub_seg);

}  // namespace rn
