// Throughput probe for the softmax design: MUFU ex2 (f32 vs f16x2) and FMA (scalar vs packed f32x2) on sm_100a.
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#define ITERS 4096
template <int MODE>
__global__ void k(float* out) {
  float a0 = threadIdx.x * 1e-3f, a1 = a0 + 1.f, a2 = a0 + 2.f, a3 = a0 + 3.f;
  unsigned h0 = 0x30003000u + threadIdx.x, h1 = h0 + 1, h2 = h0 + 2, h3 = h0 + 3;
  unsigned long long p0 = 0x3f8000003f800000ull + threadIdx.x, p1 = p0 + 1, p2 = p0 + 2, p3 = p0 + 3;
  const unsigned long long c = 0x3f0000003f000000ull;
#pragma unroll 4
  for (int i = 0; i < ITERS; ++i) {
    if (MODE == 0) {
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a0));
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a1));
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a2));
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a3));
    } else if (MODE == 1) {
      asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h0));
      asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h1));
      asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h2));
      asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h3));
    } else if (MODE == 2) {
      asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(a0) : "f"(0.5f));
      asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(a1) : "f"(0.5f));
      asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(a2) : "f"(0.5f));
      asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(a3) : "f"(0.5f));
    } else if (MODE == 3) {
      asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p0) : "l"(c));
      asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p1) : "l"(c));
      asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p2) : "l"(c));
      asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p3) : "l"(c));
    } else if (MODE == 4) {  // interleaved MUFU + FMA: do they overlap?
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a0));
      asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(a1) : "f"(0.5f));
      asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(a2) : "f"(0.5f));
      asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(a3) : "f"(0.5f));
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + __uint_as_float(h0 ^ h1 ^ h2 ^ h3) + (float)(p0 ^ p1 ^ p2 ^ p3);
}
template <int MODE>
void run(const char* name, float* d, double per_instr_results) {
  cudaEvent_t s, e;
  cudaEventCreate(&s); cudaEventCreate(&e);
  const int blocks = 148 * 4, threads = 256;
  k<MODE><<<blocks, threads>>>(d);
  cudaEventRecord(s);
  k<MODE><<<blocks, threads>>>(d);
  cudaEventRecord(e);
  cudaEventSynchronize(e);
  float ms; cudaEventElapsedTime(&ms, s, e);
  double instr = (double)blocks * threads * ITERS * 4;
  printf("%-28s %8.3f ms  %8.1f G thread-instr/s  %8.1f G results/s  (%.2f thread-instr/clk/SM @1.9GHz)\n", name, ms,
         instr / ms / 1e6, instr * per_instr_results / ms / 1e6, instr / (ms * 1e-3) / 148 / 1.9e9);
}
int main() {
  float* d; cudaMalloc(&d, 148 * 4 * 256 * 4);
  run<0>("ex2.approx.f32", d, 1);
  run<1>("ex2.approx.f16x2", d, 2);
  run<2>("fma.f32", d, 1);
  run<3>("fma.f32x2", d, 2);
  run<4>("1 ex2 + 3 fma interleaved", d, 1);
  return 0;
}
