// pipe_bench3.cu -- issue/pipe cost model for the NTT butterfly on sm_100a: how long does IMAD.WIDE.U32 hold
// the fma pipe next to IMAD / IADD3 / IADD3.X / DFMA, and which of those overlap with it.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef unsigned long long u64; typedef unsigned int u32;
#define ITERS 2048
#define W(i)    asm volatile("{ .reg .u32 lo, hi; mov.b64 {lo, hi}, %0; mul.wide.u32 %0, lo, hi; }" : "+l"(acc[i]))
#define WZ(i)   asm volatile("{ .reg .u32 lo, hi; mov.b64 {lo, hi}, %0; mul.wide.u32 %0, lo, hi; }" : "+l"(acc[i]))
#define IM(i)   asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(b), "r"(a))
#define IA(i)   asm volatile("add.u32 %0, %0, %1;" : "+r"(q[i]) : "r"(b))
#define LO(i)   asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(q[i]) : "r"(b), "r"(a))
#define SH(i)   asm volatile("shf.l.wrap.b32 %0, %0, %1, 3;" : "+r"(q[i]) : "r"(b))
#define A64(i)  asm volatile("add.cc.u32 %0, %0, %2; addc.u32 %1, %1, %3;" : "+r"(q[i]), "+r"(q[(i + 4) & 7]) : "r"(a), "r"(b))
#define DF(i)   asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"(db), "d"(da))
template <int MODE> __global__ void k(u64 *out, u32 a0, u32 b0) {
  u32 a = a0 + threadIdx.x, b = b0 | 1;
  u64 acc[8]; u32 r[8], q[8]; double d[8]; double da = a * 1e-9, db = 1.0 + b * 1e-9;
#pragma unroll
  for (int i = 0; i < 8; i++) { acc[i] = a * (i + 1); r[i] = a + i; q[i] = a ^ i; d[i] = i + a; }
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      if (MODE == 0) W(i);
      if (MODE == 1) WZ(i);
      if (MODE == 2) { W(i); IA(i); }
      if (MODE == 3) { W(i); IA(i); LO(i); }
      if (MODE == 4) { W(i); IA(i); LO(i); SH(i); }
      if (MODE == 5) { W(i); IM(i); }
      if (MODE == 6) DF(i);
      if (MODE == 7) { W(i); DF(i); }
      if (MODE == 8) { W(i); DF(i); IA(i); LO(i); }
      if (MODE == 9) { W(i); A64(i); }
      if (MODE == 10) { W(i); A64(i); IA((i + 1) & 7); }
      if (MODE == 11) { IM(i); IA(i); LO(i); }
      if (MODE == 12) { DF(i); IA(i); LO(i); SH(i); }
      if (MODE == 13) { W(i); IA(i); LO(i); SH(i); IA((i+3)&7); LO((i+5)&7); }
    }
  }
  u64 s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) s += acc[i] + r[i] + q[i] + (u64)d[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> void run(const char *name, int groups_per_iter) {
  u64 *out; cudaMalloc(&out, 148 * 1024 * 8 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int threads : {512, 1024}) {
    k<MODE><<<148 * 2, threads>>>(out, 3, 5);
    cudaEventRecord(e0);
    k<MODE><<<148 * 2, threads>>>(out, 3, 5);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    double grp = 148.0 * 2 * threads * ITERS * 8.0 * groups_per_iter;
    double cycles = ms * 1e-3 * clk * 1e3;
    // clocks per warp-group per SM sub-partition (4 per SM)
    printf("%-44s threads=%4d  %.2f groups/clk/SM = %.2f clk per warp-group per SMSP\n", name, threads, grp / cycles / 148.0, 128.0 / (grp / cycles / 148.0));
  }
  cudaFree(out);
}
int main() {
    run<1>("W (mul.wide, RZ addend)", 1);
  run<2>("W + IADD", 1);
  run<3>("W + IADD + LOP3", 1);
  run<4>("W + IADD + LOP3 + SHF", 1);
  run<13>("W + 5 ALU", 1);
  run<5>("W + IMAD", 1);
  run<6>("DFMA", 1);
  run<7>("W + DFMA", 1);
  run<8>("W + DFMA + IADD + LOP3", 1);
  run<9>("W + add64(cc)", 1);
  run<10>("W + add64(cc) + IADD", 1);
  run<11>("IMAD + IADD + LOP3", 1);
  run<12>("DFMA + IADD + LOP3 + SHF", 1);
  return 0;
}
