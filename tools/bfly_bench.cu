// bfly_bench.cu -- register-only butterfly throughput on sm_100a: how many lazy Shoup
// butterflies per clock per SM the arithmetic itself sustains (no memory traffic).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../eva_b200/csrc/modarith.cuh"
#define ITERS 512
// variant 0: exact mulhi (compiler), variant 1: 3-partial-product approximate mulhi (result < 4p)
template <int V> __device__ __forceinline__ u64 mulw(u64 y, u64 w, u64 ws, u64 np) {
  if (V == 0) { u64 q = __umul64hi(ws, y); return w * y + q * np; }
  u32 yl = (u32)y, yh = (u32)(y >> 32), sl = (u32)ws, sh = (u32)(ws >> 32);
  u64 m1 = (u64)sh * yl, m2 = (u64)sl * yh;
  u64 q = (u64)sh * yh + (m1 >> 32) + (m2 >> 32);
  return w * y + q * np;
}
template <int V, int E> __global__ void __launch_bounds__(1024, 1) k(u64 *out, u64 p, u64 w0, u64 ws0) {
  u64 x[E];
  const u64 np = 0 - p, two_p = 2 * p;
#pragma unroll
  for (int i = 0; i < E; i++) x[i] = (threadIdx.x * 977 + i * 131 + blockIdx.x) % p;
  u64 w = w0 + threadIdx.x % 7, ws = ws0 + threadIdx.x % 5;
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int d = E / 2; d >= 1; d >>= 1) {
#pragma unroll
      for (int g = 0; g < E / 2 / d; g++)
#pragma unroll
        for (int j = 0; j < d; j++) {
          const int a = g * 2 * d + j, b = a + d;
          u64 t = mulw<V>(x[b], w, ws, np);
          u64 xa = x[a];
          x[a] = xa + t; x[b] = xa - t + two_p;
        }
    }
#pragma unroll
    for (int i = 0; i < E; i++) x[i] = csub(x[i], 8 * p);  // keep values bounded (1 csub / element / log2(E) stages)
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
  printf("%-40s E=%2d threads=%4d  %.2f butterflies/clk/SM  (%.3f ms)\n", name, E, threads, bf / cycles / 148.0, ms);
  cudaFree(out);
}
int main() {
  run<0, 16>("exact mulhi", 1024); run<0, 16>("exact mulhi", 512); run<0, 32>("exact mulhi", 512);
  run<1, 16>("3-product approx mulhi", 1024); run<1, 16>("3-product approx mulhi", 512); run<1, 32>("3-product approx mulhi", 512);
  return 0;
}
