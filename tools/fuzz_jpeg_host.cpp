// ASAN/UBSAN mutation fuzz of the host JPEG parser + Huffman decoder (untrusted input):
//   g++ -std=c++17 -O1 -g -fsanitize=address,undefined -fno-sanitize-recover=undefined -Iroomnet_b200/csrc \
//       tools/fuzz_jpeg_host.cpp roomnet_b200/csrc/jpeg_host.cpp -o /tmp/fuzz_jpeg && /tmp/fuzz_jpeg seed1.jpg seed2.jpg ...
// 60,000 mutated files (byte noise, header noise, truncation, bit flips) must end in a status, never in a report.
#include "jpeg_host.h"
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <random>
#include <cstring>
int main(int argc, char** argv) {
  std::vector<std::vector<uint8_t>> seeds;
  for (int i = 1; i < argc; ++i) { FILE* f = fopen(argv[i], "rb"); std::vector<uint8_t> d; int c; while ((c = fgetc(f)) != EOF) d.push_back(c); fclose(f); seeds.push_back(d); }
  std::mt19937 rng(1);
  long ok = 0, un = 0, co = 0;
  for (int it = 0; it < 60000; ++it) {
    std::vector<uint8_t> d = seeds[it % seeds.size()];
    int nm = 1 + rng() % 8;
    int mode = rng() % 4;
    for (int k = 0; k < nm; ++k) {
      size_t p = rng() % d.size();
      if (mode == 0) d[p] = rng();
      else if (mode == 1) { if (p < 700) d[p] = rng(); else d[rng() % 700 % d.size()] = rng(); }  // headers
      else if (mode == 2) d.resize(p + 1);
      else d[p] ^= 1u << (rng() % 8);
    }
    // exact-size heap copy so that ASAN sees any over-read
    uint8_t* buf = (uint8_t*)malloc(d.size()); memcpy(buf, d.data(), d.size());
    rn::JpegInfo info;
    int st = rn::JpegParseHeader(buf, d.size(), &info);
    if (st == 0) {
      if (info.coef_count > (1u << 26)) { free(buf); continue; }
      int16_t* c = (int16_t*)malloc(info.coef_count * 2 + 2);
      st = rn::JpegDecodeCoefficients(buf, d.size(), info, c);
      free(c);
    }
    (st == 0 ? ok : st == 1 ? un : co)++;
    free(buf);
  }
  printf("ok %ld unsupported %ld corrupt %ld\n", ok, un, co);
}
