// Issue-rate micro-benchmarks for the epilogue instruction mix on sm_100a: FADD vs packed FADD2 (add.f32x2),
// FFMA vs FFMA2, HFMA2, FADD.SAT, SHFL, F2FP, half->float converts.  One CTA per SM, `nw` warps, long unrolled
// dependent chains (8 independent accumulators per thread).  Prints instructions per cycle per SM.
#include <cstdio>
#include <cstdint>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#define N_IT 4096

template <int OP>
__global__ void k(float* out, long long* cyc, float seed) {
  float a[8];
  unsigned long long p[8];
  uint32_t h[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    a[i] = seed + i + threadIdx.x;
    p[i] = (static_cast<unsigned long long>(__float_as_uint(a[i])) << 32) | __float_as_uint(a[i] * 0.5f);
    h[i] = 0x3c003c00u + i;
  }
  const float b = seed * 0.999f;
  const unsigned long long pb = (static_cast<unsigned long long>(__float_as_uint(b)) << 32) | __float_as_uint(b);
  const uint32_t hb = 0x3bff3bffu;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < N_IT; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (OP == 0) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b));
      if (OP == 1) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(pb));
      if (OP == 2) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(a[i]) : "f"(b));
      if (OP == 3) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p[i]) : "l"(pb));
      if (OP == 4) asm volatile("fma.rn.f16x2 %0, %0, %1, %1;" : "+r"(h[i]) : "r"(hb));
      if (OP == 5) asm volatile("add.sat.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b));
      if (OP == 6) asm volatile("shfl.sync.down.b32 %0, %0, 1, 0x1f, 0xffffffff;" : "+r"(h[i]));
      if (OP == 7) asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h[i]) : "f"(a[i]), "f"(a[(i + 1) & 7]));
      if (OP == 8) asm volatile("{.reg .f16 lo, hi; mov.b32 {lo, hi}, %1; cvt.f32.f16 %0, lo;}" : "=f"(a[i]) : "r"(h[i]));
      if (OP == 9) asm volatile("add.rn.f16x2 %0, %0, %1;" : "+r"(h[i]) : "r"(hb));
      if (OP == 10) {  // mix: FADD + HFMA2 alternating (fma pipe + ?)
        asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b));
        asm volatile("shfl.sync.down.b32 %0, %0, 1, 0x1f, 0xffffffff;" : "+r"(h[i]));
      }
      if (OP == 11) {
        asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b));
        asm volatile("lop3.b32 %0, %0, %1, %1, 0x96;" : "+r"(h[i]) : "r"(hb));
      }
      if (OP == 12) asm volatile("min.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b));
      if (OP == 13) asm volatile("min.f16x2 %0, %0, %1;" : "+r"(h[i]) : "r"(hb));
      if (OP == 14) asm volatile("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(h[i]) : "f"(a[i]), "f"(a[(i + 1) & 7]));
      if (OP == 15) asm volatile("{.reg .f16 lo, hi; mov.b32 {lo, hi}, %1; fma.rn.f32.f16 %0, hi, lo, %0;}" : "+f"(a[i]) : "r"(hb));
      if (OP == 16) asm volatile("{.reg .f16 lo, hi; mov.b32 {lo, hi}, %1; add.rn.f32.f16 %0, hi, %0;}" : "+f"(a[i]) : "r"(hb));
      if (OP == 17) asm volatile("prmt.b32 %0, %0, %1, 0x5410;" : "+r"(h[i]) : "r"(hb));
      if (OP == 18) asm volatile("lop3.b32 %0, %0, %1, %1, 0x96;" : "+r"(h[i]) : "r"(hb));
      if (OP == 19) {  // FADD + HADD2: one pipe or two?
        asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b));
        asm volatile("add.rn.f16x2 %0, %0, %1;" : "+r"(h[i]) : "r"(hb));
      }
      if (OP == 20) {  // FADD2 + HADD2
        asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(pb));
        asm volatile("add.rn.f16x2 %0, %0, %1;" : "+r"(h[i]) : "r"(hb));
      }
      if (OP == 21) {  // FADD + FMNMX
        asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b));
        asm volatile("min.f32 %0, %0, %1;" : "+f"(a[(i + 4) & 7]) : "f"(b));
      }
      if (OP == 22) {  // HADD2 + HMNMX2
        asm volatile("add.rn.f16x2 %0, %0, %1;" : "+r"(h[i]) : "r"(hb));
        asm volatile("min.f16x2 %0, %0, %1;" : "+r"(h[(i + 4) & 7]) : "r"(hb));
      }
      if (OP == 23) {  // HADD2 + F2FP
        asm volatile("add.rn.f16x2 %0, %0, %1;" : "+r"(h[i]) : "r"(hb));
        asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h[(i + 4) & 7]) : "f"(a[i]), "f"(a[(i + 1) & 7]));
      }
      if (OP == 24) {  // FADD + F2FP
        asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b));
        asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h[i]) : "f"(a[(i + 3) & 7]), "f"(a[(i + 1) & 7]));
      }
      if (OP == 25) {  // HADD2 + SHFL
        asm volatile("add.rn.f16x2 %0, %0, %1;" : "+r"(h[i]) : "r"(hb));
        asm volatile("shfl.sync.down.b32 %0, %0, 1, 0x1f, 0xffffffff;" : "+r"(h[(i + 4) & 7]));
      }
      if (OP == 26) {  // FHFMA + HADD2
        asm volatile("{.reg .f16 lo, hi; mov.b32 {lo, hi}, %1; fma.rn.f32.f16 %0, hi, lo, %0;}" : "+f"(a[i]) : "r"(hb));
        asm volatile("add.rn.f16x2 %0, %0, %1;" : "+r"(h[i]) : "r"(hb));
      }
      if (OP == 27) {  // F2FP + SHFL + HMNMX2 (none on the fma pipe?)
        asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h[i]) : "f"(a[(i + 3) & 7]), "f"(a[(i + 1) & 7]));
        asm volatile("min.f16x2 %0, %0, %1;" : "+r"(h[(i + 4) & 7]) : "r"(hb));
      }
      if (OP == 28) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(p[(i + 4) & 7]));
      if (OP == 29) {  // FADD2 + FADD
        asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(pb));
        asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b));
      }
      if (OP == 30) {  // HFMA2 + FMNMX + F2FP : three pipes?
        asm volatile("fma.rn.f16x2 %0, %0, %1, %1;" : "+r"(h[i]) : "r"(hb));
        asm volatile("min.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b));
      }
    }
  }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i] + __uint_as_float(static_cast<uint32_t>(p[i])) + __uint_as_float(h[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char* name, int nw, int per_it) {
  float* out;
  long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4);
  cudaMalloc(&cyc, 148 * 8);
  k<OP><<<148, nw * 32>>>(out, cyc, 1.0f);
  k<OP><<<148, nw * 32>>>(out, cyc, 1.0f);
  cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double c = 0;
  for (int i = 0; i < 148; ++i) c += h[i];
  c /= 148;
  double inst = double(N_IT) * 8 * per_it * nw;
  printf("%-28s warps %2d: %.3f warp-instr/cycle/SM (%.3f per SMSP)\n", name, nw, inst / c, inst / c / 4);
  cudaFree(out);
  cudaFree(cyc);
}

int main() {
  for (int nw : {4, 16}) {
    run<0>("FADD", nw, 1);
    run<1>("FADD2 (add.f32x2)", nw, 1);
    run<2>("FFMA", nw, 1);
    run<3>("FFMA2 (fma.f32x2)", nw, 1);
    run<4>("HFMA2", nw, 1);
    run<9>("HADD2", nw, 1);
    run<5>("FADD.SAT", nw, 1);
    run<6>("SHFL.DOWN", nw, 1);
    run<7>("F2FP (cvt f16x2.f32)", nw, 1);
    run<8>("HADD2.F32 (cvt f32.f16)", nw, 1);
    run<10>("FADD + SHFL mix", nw, 2);
    run<11>("FADD + LOP3 mix", nw, 2);
    run<12>("FMNMX", nw, 1);
    run<13>("HMNMX2", nw, 1);
    run<14>("F2FP.RELU", nw, 1);
    run<15>("FHFMA (fma.f32.f16)", nw, 1);
    run<16>("FHADD (add.f32.f16)", nw, 1);
    run<17>("PRMT", nw, 1);
    run<18>("LOP3", nw, 1);
    run<19>("FADD + HADD2 mix", nw, 2);
    run<20>("FADD2 + HADD2 mix", nw, 2);
    run<21>("FADD + FMNMX mix", nw, 2);
    run<22>("HADD2 + HMNMX2 mix", nw, 2);
    run<23>("HADD2 + F2FP mix", nw, 2);
    run<24>("FADD + F2FP mix", nw, 2);
    run<25>("HADD2 + SHFL mix", nw, 2);
    run<26>("FHFMA + HADD2 mix", nw, 2);
    run<27>("F2FP + HMNMX2 mix", nw, 2);
    run<28>("FADD2 (reg,reg)", nw, 1);
    run<29>("FADD2 + FADD mix", nw, 2);
    run<30>("HFMA2 + FMNMX mix", nw, 2);
  }
  return 0;
}
