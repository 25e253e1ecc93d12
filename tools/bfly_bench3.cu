// bfly_bench3.cu -- "two-row fold" modular multiply for primes p = 2^60 - delta (SEAL's 60-bit NTT primes):
//   v = w * 2^32 mod p (precomputed beside w), eps = 2^61 mod p = 2*delta (< 2^26)
//   S = y0*w + y1*v  (< 2^33 p, == y*w mod p, 4 IMAD.WIDE)      y = y1*2^32 + y0, ANY u64
//   r = (S mod 2^61) + (S >> 61) * eps   (1 IMAD.WIDE)           r == y*w (mod p), r < 2^61 + 2^58 < 2.26p
// Register-only butterfly throughput vs the lazy Shoup butterfly of round 1 (tools/bfly_bench2.cu).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef uint64_t u64; typedef uint32_t u32;
#define ITERS 512

__device__ __forceinline__ u64 madw(u32 a, u32 b, u64 c) { u64 d; asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(d) : "r"(a), "r"(b), "l"(c)); return d; }
__device__ __forceinline__ u64 mulw(u32 a, u32 b) { u64 d; asm("mul.wide.u32 %0, %1, %2;" : "=l"(d) : "r"(a), "r"(b)); return d; }

// variant A: C with __int128
__device__ __forceinline__ u64 fold_mul_A(u64 y, u64 w, u64 v, u32 eps) {
  const u32 y0 = (u32)y, y1 = (u32)(y >> 32);
  const u32 w0 = (u32)w, w1 = (u32)(w >> 32), v0 = (u32)v, v1 = (u32)(v >> 32);
  u64 B = madw(y1, v1, mulw(y0, w1));
  unsigned __int128 S = (unsigned __int128)mulw(y0, w0) + mulw(y1, v0) + ((unsigned __int128)B << 32);
  const u32 H = (u32)(S >> 61);
  const u64 L = (u64)S & ((1ull << 61) - 1);
  return madw(H, eps, L);
}
// variant B: explicit carry chain in PTX
__device__ __forceinline__ u64 fold_mul_B(u64 y, u64 w, u64 v, u32 eps) {
  const u32 y0 = (u32)y, y1 = (u32)(y >> 32);
  const u32 w0 = (u32)w, w1 = (u32)(w >> 32), v0 = (u32)v, v1 = (u32)(v >> 32);
  const u64 B = madw(y1, v1, mulw(y0, w1));
  const u64 A = mulw(y0, w0), A2 = mulw(y1, v0);
  u32 s0, s1, s2;
  asm("{\n\t"
      ".reg .u32 t;\n\t"
      "add.cc.u32 %0, %3, %5;\n\t"
      "addc.cc.u32 t, %4, %6;\n\t"
      "addc.u32 %2, %8, 0;\n\t"
      "add.cc.u32 %1, t, %7;\n\t"
      "addc.u32 %2, %2, 0;\n\t"
      "}" : "=r"(s0), "=r"(s1), "=r"(s2)
      : "r"((u32)A), "r"((u32)(A >> 32)), "r"((u32)A2), "r"((u32)(A2 >> 32)), "r"((u32)B), "r"((u32)(B >> 32)));
  u32 H; asm("shf.l.wrap.b32 %0, %1, %2, 3;" : "=r"(H) : "r"(s1), "r"(s2));
  const u64 L = ((u64)(s1 & 0x1fffffffu) << 32) | s0;
  return madw(H, eps, L);
}
// round-1 lazy Shoup product (exact quotient)
__device__ __forceinline__ u64 shoup_lazy_n(u64 y, u64 w, u64 ws, u64 np) { return w * y + __umul64hi(ws, y) * np; }

template <int V, int E> __global__ void __launch_bounds__(1024, 1) k(u64 *out, u64 p, u64 w0, u64 ws0, u32 eps) {
  u64 x[E];
  const u64 np = 0 - p, two_p = 2 * p, three_p = 3 * p;
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
          const u64 xa = x[a];
          if (V == 0) { u64 t = shoup_lazy_n(x[b], w, ws, np); x[a] = xa + t; x[b] = xa - t + two_p; }
          if (V == 1) { u64 t = fold_mul_A(x[b], w, ws, eps); x[a] = xa + t; x[b] = xa + three_p - t; }
          if (V == 2) { u64 t = fold_mul_B(x[b], w, ws, eps); x[a] = xa + t; x[b] = xa + three_p - t; }
        }
    }
    // one bound fix per 4 stages on every element (upper bound of what the transform needs)
#pragma unroll
    for (int i = 0; i < E; i++) {
      if (V == 0) x[i] = x[i] >= 8 * p ? x[i] - 8 * p : x[i];
      else x[i] = madw((u32)(x[i] >> 61), eps, x[i] & ((1ull << 61) - 1));
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
  const u32 eps = (u32)((1ull << 61) % p);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<V, E><<<148, threads>>>(out, p, w, ws, eps);
  cudaEventRecord(e0);
  k<V, E><<<148, threads>>>(out, p, w, ws, eps);
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
  run<0, 16>("V0 lazy Shoup (round 1), csub 8p per 4 stages", 1024);
  run<1, 16>("V1 two-row fold, __int128 sum", 1024);
  run<2, 16>("V2 two-row fold, PTX carry chain", 1024);
  run<1, 16>("V1 512 thr", 512);
  run<2, 16>("V2 512 thr", 512);
  run<2, 8>("V2 E=8", 1024);
  return 0;
}
