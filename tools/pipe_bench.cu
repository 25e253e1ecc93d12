// pipe_bench.cu -- integer pipe throughput microbenchmark for sm_100a (run under gpurun):
// how many IMAD / IMAD.WIDE / IMAD.HI / IADD3 / 64-bit mul-hi instructions per clock per SM.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef unsigned long long u64; typedef unsigned int u32;
#define ITERS 4096
template <int MODE> __global__ void k(u64 *out, u32 a0, u32 b0) {
  u32 a = a0 + threadIdx.x, b = b0 | 1;
  u64 acc[8]; u32 r[8];
#pragma unroll
  for (int i = 0; i < 8; i++) { acc[i] = a * (i + 1); r[i] = a + i; }
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      if (MODE == 0) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(b), "r"(a));           // IMAD
      if (MODE == 1) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[i]) : "r"(r[i]), "r"(b));    // IMAD.WIDE.U32
      if (MODE == 2) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(b), "r"(a));           // IMAD.HI.U32
      if (MODE == 3) asm volatile("add.u32 %0, %0, %1;" : "+r"(r[i]) : "r"(b));                          // IADD3
      if (MODE == 4) asm volatile("mul.hi.u64 %0, %0, %1;" : "+l"(acc[i]) : "l"((u64)b << 20 | a));     // 64-bit mulhi sequence
      if (MODE == 5) asm volatile("mul.lo.u64 %0, %0, %1;" : "+l"(acc[i]) : "l"((u64)b << 20 | a));     // 64-bit mullo sequence
      if (MODE == 6) asm volatile("add.cc.u32 %0, %0, %2; addc.u32 %1, %1, %3;" : "+r"(r[i]), "+r"(r[(i + 1) & 7]) : "r"(a), "r"(b)); // 64-bit add
      if (MODE == 7) { asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(b), "r"(a)); asm volatile("add.u32 %0, %0, %1;" : "+r"(r[(i + 4) & 7]) : "r"(b)); } // IMAD + IADD3 dual
    }
  }
  u64 s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) s += acc[i] + r[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> void run(const char *name, int instr_per_iter) {
  u64 *out; cudaMalloc(&out, 148 * 1024 * 8 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int threads : {256, 1024}) {
    k<MODE><<<148 * 2, threads>>>(out, 3, 5);
    cudaEventRecord(e0);
    k<MODE><<<148 * 2, threads>>>(out, 3, 5);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    double instr = 148.0 * 2 * threads * ITERS * 8.0 * instr_per_iter;
    double cycles = ms * 1e-3 * clk * 1e3;
    printf("%-28s threads=%4d  %.1f thread-instr/clk/SM  (%.3f ms, clock %d kHz)\n", name, threads, instr / cycles / 148.0, ms, clk);
  }
  cudaFree(out);
}
int main() {
  run<0>("IMAD (mad.lo.u32)", 1);
  run<1>("IMAD.WIDE.U32", 1);
  run<2>("IMAD.HI.U32", 1);
  run<3>("IADD3 (add.u32)", 1);
  run<4>("mul.hi.u64 (sequence)", 1);
  run<5>("mul.lo.u64 (sequence)", 1);
  run<6>("64-bit add (2 instr)", 2);
  run<7>("IMAD + IADD3 interleaved", 2);
  return 0;
}
