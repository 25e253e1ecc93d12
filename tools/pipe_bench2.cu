// pipe_bench2.cu -- throughput of carry-chain integer instructions on sm_100a
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64; typedef unsigned int u32;
#define ITERS 4096
template <int MODE> __global__ void k(u64 *out, u32 a0, u32 b0) {
  u32 a = a0 + threadIdx.x, b = b0 | 1;
  u32 lo[8], hi[8];
#pragma unroll
  for (int i = 0; i < 8; i++) { lo[i] = a * (i + 1); hi[i] = a + i; }
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      if (MODE == 0) asm volatile("add.cc.u32 %0, %0, %2; addc.u32 %1, %1, %3;" : "+r"(lo[i]), "+r"(hi[i]) : "r"(a), "r"(b));              // IADD3 + IADD3.X
      if (MODE == 1) asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.u32 %1, %2, %3, %1;" : "+r"(lo[i]), "+r"(hi[i]) : "r"(a), "r"(b)); // IMAD + IMAD.HI.X ?
      if (MODE == 2) asm volatile("add.cc.u32 %0, %0, %2; madc.lo.u32 %1, %2, %3, %1;" : "+r"(lo[i]), "+r"(hi[i]) : "r"(a), "r"(b));        // IADD3 + IMAD.X
      if (MODE == 3) { u64 v = ((u64)hi[i] << 32) | lo[i]; u64 c = ((u64)b << 32) | a; v = v >= c ? v - c : v; v += c >> 1; lo[i] = (u32)v; hi[i] = (u32)(v >> 32); } // 64-bit csub + add
      if (MODE == 4) { u64 v = ((u64)hi[i] << 32) | lo[i]; u64 c = ((u64)b << 32) | a; v = v + c; lo[i] = (u32)v; hi[i] = (u32)(v >> 32); }      // plain 64-bit add via compiler
      if (MODE == 5) { u64 v = ((u64)hi[i] << 32) | lo[i]; u64 c = ((u64)b << 32) | a; v = v * c; lo[i] = (u32)v; hi[i] = (u32)(v >> 32); }      // mul.lo.u64
      if (MODE == 6) { u64 v = ((u64)hi[i] << 32) | lo[i]; u64 c = ((u64)b << 32) | a; v = __umul64hi(v, c) + 3; lo[i] = (u32)v; hi[i] = (u32)(v >> 32); } // mul.hi.u64
    }
  }
  u64 s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) s += lo[i] + hi[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> void run(const char *name, int ops) {
  u64 *out; cudaMalloc(&out, 148 * 2 * 1024 * 8);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<148 * 2, 1024>>>(out, 3, 5);
  cudaEventRecord(e0);
  k<MODE><<<148 * 2, 1024>>>(out, 3, 5);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  double n = 148.0 * 2 * 1024 * ITERS * 8.0 * ops;
  double cycles = ms * 1e-3 * clk * 1e3;
  printf("%-44s %.1f ops/clk/SM (%.3f ms)\n", name, n / cycles / 148.0, ms);
  cudaFree(out);
}
int main() {
  run<0>("add.cc + addc  (64-bit add), per pair", 1);
  run<1>("mad.lo.cc + madc.hi, per pair", 1);
  run<2>("add.cc + madc.lo, per pair", 1);
  run<3>("64-bit csub + add (compiler), per op", 1);
  run<4>("64-bit add (compiler), per op", 1);
  run<5>("mul.lo.u64 (compiler), per op", 1);
  run<6>("mul.hi.u64 (compiler), per op", 1);
  return 0;
}
