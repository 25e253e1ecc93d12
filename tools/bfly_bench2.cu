// bfly_bench2.cu -- formulations of the lazy Shoup butterfly, register-only throughput (sm_100a)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../eva_b200/csrc/modarith.cuh"
#define ITERS 512
__device__ __forceinline__ u64 mulhi_4wide(u64 a, u64 b) {  // schoolbook, no carry-in multiplies
  u32 a0 = (u32)a, a1 = (u32)(a >> 32), b0 = (u32)b, b1 = (u32)(b >> 32);
  u64 t0, t1, t2, t3;
  asm("mul.wide.u32 %0, %1, %2;" : "=l"(t0) : "r"(a0), "r"(b0));
  asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(t1) : "r"(a1), "r"(b0), "l"(t0 >> 32));
  asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(t2) : "r"(a0), "r"(b1), "l"(t1 & 0xffffffffull));
  asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(t3) : "r"(a1), "r"(b1), "l"(t1 >> 32));
  return t3 + (t2 >> 32);
}
__device__ __forceinline__ u64 mulhi_approx(u64 a, u64 b) {  // >= true - 2: drops a0*b0 and the low halves of the cross terms
  u32 a0 = (u32)a, a1 = (u32)(a >> 32), b0 = (u32)b, b1 = (u32)(b >> 32);
  u32 c0, c1; u64 t;
  asm("mul.hi.u32 %0, %1, %2;" : "=r"(c0) : "r"(a1), "r"(b0));
  asm("mul.hi.u32 %0, %1, %2;" : "=r"(c1) : "r"(a0), "r"(b1));
  asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(t) : "r"(a1), "r"(b1), "l"((u64)c0));
  return t + c1;
}
// hand-scheduled forward butterfly: s = xa + w*y + q*np with q = mulhi_approx(ws, y); returns s,
// xb = 2*xa + bias - s.  3 mad.wide + 2 mul.hi + 4 mad.lo on the fma pipe.
__device__ __forceinline__ void bfly_ptx(u64 &xa, u64 &xb, u64 w, u64 ws, u64 np, u64 bias) {
  u32 y0 = (u32)xb, y1 = (u32)(xb >> 32), w0 = (u32)w, w1 = (u32)(w >> 32);
  u32 s0 = (u32)ws, s1 = (u32)(ws >> 32), n0 = (u32)np, n1 = (u32)(np >> 32);
  u32 c0, c1; u64 q, m;
  asm("mul.hi.u32 %0, %1, %2;" : "=r"(c0) : "r"(s1), "r"(y0));
  asm("mul.hi.u32 %0, %1, %2;" : "=r"(c1) : "r"(s0), "r"(y1));
  asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(q) : "r"(s1), "r"(y1), "l"((u64)c0));
  q += c1;
  u32 q0 = (u32)q, q1 = (u32)(q >> 32);
  asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(m) : "r"(w0), "r"(y0), "l"(xa));
  asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(m) : "r"(q0), "r"(n0), "l"(m));
  u32 h = (u32)(m >> 32);
  asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(h) : "r"(w0), "r"(y1));
  asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(h) : "r"(w1), "r"(y0));
  asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(h) : "r"(q0), "r"(n1));
  asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(h) : "r"(q1), "r"(n0));
  u64 s = ((u64)h << 32) | (u32)m;
  xb = xa + xa + bias - s;
  xa = s;
}
template <int V> __device__ __forceinline__ u64 mulw(u64 y, u64 w, u64 ws, u64 p, u64 np) {
  if (V == 0 || V == 1) { u64 q = __umul64hi(ws, y); return w * y + q * np; }
  if (V == 2) { u64 q = __umul64hi(ws, y); u64 a = w * y, b = q * p; asm volatile("" : "+l"(a), "+l"(b)); return a - b; }
  if (V == 3) { u64 q = mulhi_4wide(ws, y); return w * y + q * np; }
  if (V == 4) { u64 q = mulhi_4wide(ws, y); u64 a = w * y, b = q * p; asm volatile("" : "+l"(a), "+l"(b)); return a - b; }
  if (V == 5 || V == 7) { u64 q = mulhi_approx(ws, y); return w * y + q * np; }
  if (V == 6) { u64 q = __umul64hi(ws, y); return w * y + q * np; }
  return 0;
}
template <int V, int E> __global__ void __launch_bounds__(1024, 1) k(u64 *out, u64 p, u64 w0, u64 ws0) {
  u64 x[E];
  const u64 np = 0 - p, two_p = 2 * p;
#pragma unroll
  for (int i = 0; i < E; i++) x[i] = (u64)(threadIdx.x * 977 + i * 131 + blockIdx.x) * 0x9E3779B97F4A7C15ull >> 5;
  u64 w = w0 + threadIdx.x % 7, ws = ws0 + threadIdx.x % 5;
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int d = E / 2; d >= 1; d >>= 1) {
#pragma unroll
      for (int g = 0; g < E / 2 / d; g++)
#pragma unroll
        for (int j = 0; j < d; j++) {
          const int a = g * 2 * d + j, b = a + d;
          u64 xa = x[a];
          if (V == 8) { bfly_ptx(x[a], x[b], w, ws, np, 2 * two_p); }
          else if (V >= 6) {
            u64 s = xa + mulw<V>(x[b], w, ws, p, np);
            x[b] = (xa + xa + (V == 7 ? 2 * two_p : two_p)) - s; x[a] = s;
          } else {
          u64 t = mulw<V>(x[b], w, ws, p, np);
          x[a] = xa + t; x[b] = xa - t + (V == 5 ? 2 * two_p : two_p);
          }
        }
    }
    if (V != 1) {
#pragma unroll
      for (int i = 0; i < E; i++) x[i] = csub(x[i], 8 * p);
    }
  }
  u64 s = 0;
#pragma unroll
  for (int i = 0; i < E; i++) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int V, int E> void run(const char *name, int threads) {
  u64 *out; cudaMalloc(&out, 148 * 1024 * 8);
  const u64 p = 0xffffffffffc0001ull, w = 0x123456789abcdefull, ws = 0x2468acf13579bdfull;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<V, E><<<148, threads>>>(out, p, w, ws);
  cudaEventRecord(e0);
  k<V, E><<<148, threads>>>(out, p, w, ws);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  int logE = 0; while ((1 << logE) < E) logE++;
  double bf = 148.0 * threads * ITERS * (E / 2) * logE;
  double cycles = ms * 1e-3 * clk * 1e3;
  printf("%-52s E=%2d thr=%4d  %.2f bf/clk/SM\n", name, E, threads, bf / cycles / 148.0);
  cudaFree(out);
}
int main() {
  run<0, 16>("V0 exact mulhi (compiler), fused low products", 1024);
  run<1, 16>("V1 = V0 without the per-iteration csub", 1024);
  run<2, 16>("V2 compiler mulhi, separate low products", 1024);
  run<3, 16>("V3 4x mad.wide mulhi (no .X), fused low", 1024);
  run<4, 16>("V4 4x mad.wide mulhi, separate low", 1024);
  run<3, 8>("V3 E=8", 1024);
  run<6, 16>("V6 exact mulhi, add folded into the mad chain", 1024);
  run<7, 16>("V7 approx mulhi, add folded", 1024);
  run<8, 16>("V8 hand-written PTX butterfly (approx mulhi, folded)", 1024);
  run<5, 16>("V5 approx mulhi (2 mul.hi + 1 mad.wide), fused low", 1024);
  return 0;
}
