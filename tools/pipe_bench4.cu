// pipe_bench4.cu -- per-instruction pipe occupancy on sm_100a (B200), clocks per warp-instruction per SM
// sub-partition, alone and next to IMAD.WIDE.U32.  SASS of every mode is checked with tools/sass_hist.py.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef unsigned long long u64; typedef unsigned int u32;
#define ITERS 1024
#define W(i)    asm volatile("{ .reg .u32 lo, hi; mov.b64 {lo, hi}, %0; mul.wide.u32 %0, lo, hi; }" : "+l"(acc[i]))
#define WA(i)   asm volatile("{ .reg .u32 lo, hi; mov.b64 {lo, hi}, %0; mad.wide.u32 %0, lo, hi, %1; }" : "+l"(acc[i]) : "l"(acc[(i + 1) & 7]))
#define IM(i)   asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(r[(i + 1) & 7]), "r"(r[(i + 2) & 7]))
#define IA(i)   asm volatile("add.u32 %0, %0, %1;" : "+r"(q[i]) : "r"(q[(i + 1) & 7]))
#define IA3(i)  asm volatile("{ .reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2; }" : "+r"(q[i]) : "r"(q[(i + 1) & 7]), "r"(q[(i + 2) & 7]))
#define LO(i)   asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(q[i]) : "r"(q[(i + 1) & 7]), "r"(q[(i + 3) & 7]))
#define SH(i)   asm volatile("shf.l.wrap.b32 %0, %0, %1, 3;" : "+r"(q[i]) : "r"(q[(i + 1) & 7]))
#define A64(i)  asm volatile("add.cc.u32 %0, %0, %2; addc.u32 %1, %1, %3;" : "+r"(q[i]), "+r"(s[i]) : "r"(q[(i + 1) & 7]), "r"(s[(i + 1) & 7]))
#define SEL(i)  asm volatile("{ .reg .pred p; setp.ge.u32 p, %0, %1; selp.u32 %0, %1, %2, p; }" : "+r"(q[i]) : "r"(q[(i + 1) & 7]), "r"(q[(i + 2) & 7]))
#define DF(i)   asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"(d[(i + 1) & 7]), "d"(da))
template <int MODE> __global__ void k(u64 *out, u32 a0, u32 b0) {
  u32 a = a0 + threadIdx.x, b = b0 | 1;
  u64 acc[8]; u32 r[8], q[8], s[8]; double d[8]; double da = a * 1e-9;
#pragma unroll
  for (int i = 0; i < 8; i++) { acc[i] = (u64)a * (i + 1) * 0x9E3779B97F4A7C15ull; r[i] = a + i * b; q[i] = a ^ (i * b); s[i] = a * i; d[i] = i + a; }
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      if (MODE == 0) W(i);
      if (MODE == 1) WA(i);
      if (MODE == 2) IM(i);
      if (MODE == 3) IA(i);
      if (MODE == 4) LO(i);
      if (MODE == 5) SH(i);
      if (MODE == 6) A64(i);
      if (MODE == 7) SEL(i);
      if (MODE == 8) DF(i);
      if (MODE == 9) { IA(i); LO((i + 4) & 7); }
      if (MODE == 10) { W(i); LO(i); }
      if (MODE == 11) { W(i); LO(i); LO((i + 4) & 7); }
      if (MODE == 12) { W(i); LO(i); LO((i + 4) & 7); LO((i + 2) & 7); }
      if (MODE == 13) { W(i); LO(i); LO((i + 4) & 7); LO((i + 2) & 7); LO((i + 6) & 7); }
      if (MODE == 14) { W(i); A64(i); }
      if (MODE == 15) { W(i); A64(i); A64((i + 4) & 7); }
      if (MODE == 16) { W(i); IM(i); LO(i); LO((i + 4) & 7); }
      if (MODE == 17) { W(i); DF(i); LO(i); LO((i + 4) & 7); }
      if (MODE == 18) { W(i); SH(i); SH((i + 4) & 7); }
      if (MODE == 19) { IM(i); LO(i); }
      if (MODE == 20) { IM(i); LO(i); LO((i + 4) & 7); }
      if (MODE == 21) { DF(i); LO(i); LO((i + 4) & 7); }
      if (MODE == 22) { WA(i); LO(i); LO((i + 4) & 7); }
    }
  }
  u64 t = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) t += acc[i] + r[i] + q[i] + s[i] + (u64)d[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = t;
}
template <int MODE> void run(const char *name) {
  u64 *out; cudaMalloc(&out, 148 * 1024 * 8 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int threads = 1024;
  k<MODE><<<148 * 2, threads>>>(out, 3, 5);
  cudaEventRecord(e0);
  k<MODE><<<148 * 2, threads>>>(out, 3, 5);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  double grp = 148.0 * 2 * threads * ITERS * 8.0;
  double cycles = ms * 1e-3 * clk * 1e3;
  printf("mode %2d %-40s %.2f clk per warp-group per SMSP\n", MODE, name, 128.0 / (grp / cycles / 148.0));
  cudaFree(out);
}
int main() {
  run<0>("W (mul.wide)"); run<1>("W (mad.wide, 64-bit addend)"); run<2>("IMAD"); run<3>("IADD"); run<4>("LOP3"); run<5>("SHF");
  run<6>("add.cc+addc"); run<7>("ISETP+SEL"); run<8>("DFMA"); run<9>("IADD+LOP3");
  run<10>("W + LOP3"); run<11>("W + 2 LOP3"); run<12>("W + 3 LOP3"); run<13>("W + 4 LOP3");
  run<14>("W + add64"); run<15>("W + 2 add64"); run<16>("W + IMAD + 2 LOP3"); run<17>("W + DFMA + 2 LOP3"); run<18>("W + 2 SHF");
  run<19>("IMAD + LOP3"); run<20>("IMAD + 2 LOP3"); run<21>("DFMA + 2 LOP3"); run<22>("WA + 2 LOP3");
  return 0;
}
