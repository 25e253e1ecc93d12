// bfly_bench4.cu -- instruction-count variants of the two-row fold butterfly (see bfly_bench3.cu)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef uint64_t u64; typedef uint32_t u32;
#define ITERS 512
__device__ __forceinline__ u64 madw(u32 a, u32 b, u64 c) { u64 d; asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(d) : "r"(a), "r"(b), "l"(c)); return d; }
__device__ __forceinline__ u64 mulw(u32 a, u32 b) { u64 d; asm("mul.wide.u32 %0, %1, %2;" : "=l"(d) : "r"(a), "r"(b)); return d; }
__device__ __forceinline__ u32 lo32(u64 x) { return (u32)x; }
__device__ __forceinline__ u32 hi32(u64 x) { return (u32)(x >> 32); }
__device__ __forceinline__ u64 pack(u32 lo, u32 hi) { u64 d; asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(hi)); return d; }

// S = y0*w + y1*v as three words; V selects how the sum is formed
template <int V> __device__ __forceinline__ void prod96(u64 y, u64 w, u64 v, u32 &s0, u32 &s1, u32 &s2) {
  const u32 y0 = lo32(y), y1 = hi32(y);
  const u64 B = madw(y1, hi32(v), mulw(y0, hi32(w)));
  const u64 A = mulw(y0, lo32(w)), A2 = mulw(y1, lo32(v));
  if (V == 0) {   // 5-instruction 32-bit carry chain
    asm("{\n\t.reg .u32 t;\n\t"
        "add.cc.u32 %0, %3, %5;\n\taddc.cc.u32 t, %4, %6;\n\taddc.u32 %2, %8, 0;\n\t"
        "add.cc.u32 %1, t, %7;\n\taddc.u32 %2, %2, 0;\n\t}"
        : "=r"(s0), "=r"(s1), "=r"(s2) : "r"(lo32(A)), "r"(hi32(A)), "r"(lo32(A2)), "r"(hi32(A2)), "r"(lo32(B)), "r"(hi32(B)));
  } else if (V == 1) {  // 64-bit adds in C
    const u64 T = A + A2; const u32 c = T < A;
    const u64 U = (u64)hi32(T) + lo32(B);
    s0 = lo32(T); s1 = lo32(U); s2 = hi32(B) + hi32(U) + c;
  } else {  // B.lo added into A2's high word first (A2 < 2^64 - 2^33: hi32(A2) + lo32(B) may carry) -- 64-bit chain
    u64 T, U;
    asm("{\n\t.reg .u64 bl;\n\tmov.b64 bl, {%5, %6};\n\t"   // bl = lo32(B) << 32 is not expressible; use 0:lo32(B)
        "add.cc.u64 %0, %2, %3;\n\taddc.u64 %1, %4, 0;\n\t}"
        : "=l"(T), "=l"(U) : "l"(A), "l"(A2), "l"((u64)hi32(B)), "r"(0u), "r"(0u));
    const u64 X = (u64)hi32(T) + lo32(B);
    s0 = lo32(T); s1 = lo32(X); s2 = lo32(U) + hi32(X);
  }
}
// butterfly variants: F selects how the fold and the adds are arranged
template <int V, int F> __device__ __forceinline__ void bfly(u64 &xa, u64 &xb, u64 w, u64 v, u32 eps, u64 three_p) {
  u32 s0, s1, s2; prod96<V>(xb, w, v, s0, s1, s2);
  u32 H; asm("shf.l.wrap.b32 %0, %1, %2, 3;" : "=r"(H) : "r"(s1), "r"(s2));
  const u64 L = pack(s0, s1 & 0x1fffffffu);
  if (F == 0) { const u64 t = madw(H, eps, L); const u64 X = xa; xa = X + t; xb = X + three_p - t; }
  if (F == 1) { const u64 X = xa; const u64 s = madw(H, eps, X) + L; xa = s; xb = (X + X + three_p) - s; }
  if (F == 2) { const u64 X = xa; const u64 t = mulw(H, eps) + L; xa = X + t; xb = X + three_p - t; }
}
template <int V, int F, int E> __global__ void __launch_bounds__(1024, 1) k(u64 *out, u64 p, u64 w0, u64 ws0, u32 eps) {
  u64 x[E];
  const u64 three_p = 3 * p;
#pragma unroll
  for (int i = 0; i < E; i++) x[i] = (u64)(threadIdx.x * 977 + i * 131 + blockIdx.x) * 0x9E3779B97F4A7C15ull >> 5;
  u64 w = w0 + threadIdx.x % 7, ws = ws0 + threadIdx.x % 5;
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int d = E / 2; d >= 1; d >>= 1)
#pragma unroll
      for (int g = 0; g < E / 2 / d; g++)
#pragma unroll
        for (int j = 0; j < d; j++) bfly<V, F>(x[g * 2 * d + j], x[g * 2 * d + j + d], w, ws, eps, three_p);
    // bound fix on half the elements per 4 stages (what the transform needs)
#pragma unroll
    for (int i = 0; i < E / 2; i++) x[i] = madw((u32)(x[i] >> 61), eps, x[i] & ((1ull << 61) - 1));
  }
  u64 s = 0;
#pragma unroll
  for (int i = 0; i < E; i++) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int V, int F, int E> void run(const char *name, int threads) {
  u64 *out; cudaMalloc(&out, 148 * 1024 * 8);
  const u64 p = 0xffffffffffc0001ull, w = 0x123456789abcdefull, ws = 0x2468acf13579bdfull;
  const u32 eps = (u32)((1ull << 61) % p);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<V, F, E><<<148, threads>>>(out, p, w, ws, eps);
  cudaEventRecord(e0);
  k<V, F, E><<<148, threads>>>(out, p, w, ws, eps);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  int logE = 0; while ((1 << logE) < E) logE++;
  double bf = 148.0 * threads * ITERS * (E / 2) * logE;
  double cycles = ms * 1e-3 * clk * 1e3;
  printf("V%d F%d %-40s E=%2d thr=%4d  %.2f bf/clk/SM  (%.1f clk per warp-butterfly per SMSP)\n", V, F, name, E, threads, bf / cycles / 148.0, 128.0 / (bf / cycles / 148.0));
  cudaFree(out);
}
int main() {
  run<0, 0, 16>("carry chain / fold addend L", 1024);
  run<0, 1, 16>("carry chain / fold addend X", 1024);
  run<0, 2, 16>("carry chain / separate add", 1024);
  run<1, 0, 16>("C 64-bit sum / fold addend L", 1024);
  run<1, 1, 16>("C 64-bit sum / fold addend X", 1024);
  run<2, 0, 16>("u64 cc chain / fold addend L", 1024);
  run<2, 1, 16>("u64 cc chain / fold addend X", 1024);
  return 0;
}
