// evab200.cu -- C-ABI implementation (include/evab200.h): CUDA backend for the
// op compositions in ops_impl.hpp, sm_100a kernels around the bodies in
// ntt_kernels.cuh / ops_kernels.cuh.  Product code: never includes or links
// anything under oracle/.
#include "../../include/evab200.h"
#include "host_tables.hpp"
#include "ops_impl.hpp"
#include <cuda_runtime.h>
#include <atomic>
#include <cstdlib>
#include <map>
#include <mutex>
#include <string>
#include <vector>

static thread_local std::string g_err;
static int fail(const std::string &m) { g_err = m; return 1; }
#define CUDA_OK(x)                                                                       \
  do {                                                                                   \
    cudaError_t e_ = (x);                                                                \
    if (e_ != cudaSuccess) return fail(std::string(#x) + ": " + cudaGetErrorString(e_)); \
  } while (0)

extern "C" const char *evab_last_error(void) { return g_err.c_str(); }
extern "C" int evab_version(void) { return 1; }

struct evab_ctx {
  CtxView v;
  int device, sms;
  int ntt_cluster = 0;         // CTAs per residue (0 = automatic), evab_ctx_set_ntt_cluster
  std::vector<u64> primes;
  std::vector<FoldPrime> fold_host;
  PrimeDev *d_primes = nullptr;
  u64x2 *d_tw = nullptr;
  u64x2 *d_qinv = nullptr;
  u64x2 *d_qinv_f = nullptr;
  u64 *d_halfmod = nullptr;
  u64 *d_zeros = nullptr;
  cplx *d_roots = nullptr;
  u32 *d_slot = nullptr;
  u64 *d_pow2 = nullptr;
  std::map<u64, u32 *> perms;  // galois elt -> device permutation table (NTT domain)
  std::map<u64, u32 *> cperms; // galois elt -> device signed gather table (coefficient domain)
  std::mutex mu;
  mutable std::atomic<unsigned long long> launches{0};
};
static cudaStream_t S(void *s) { return (cudaStream_t)s; }

// thread-local batched-issue state (evab_set_batch)
struct BatchState { int batch = 1; long long stride = 0, vstride = 0; };
static thread_local BatchState g_batch;
extern "C" int evab_set_ntt_cluster(int ctas_per_residue);
extern "C" int evab_set_batch(int batch, size_t stride_words, size_t value_stride) {
  if (batch < 1 || batch > 65535) return fail("evab_set_batch: batch must be in [1, 65535]");
  g_batch.batch = batch; g_batch.stride = (long long)stride_words; g_batch.vstride = (long long)value_stride;
  return 0;
}


// ---------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------
// ---- cluster protocol of the one exchange that crosses CTAs (CL > 1) ---------------------------------------
// Round 1 used barrier.cluster.arrive.release / wait.acquire around the distributed-shared-memory stores; ptxas
// turns those into MEMBAR.ALL.GPU + ERRBAR on the arrive and CCTL.IVALL (L1 invalidate) on the wait -- 7 % of
// all stall samples on the fences and 8 % on the barrier itself (profiles/r02_ntt_fold.md).  Now every CTA owns
// two mbarriers behind its exchange slice:
//   full : completes when the N/CL * 8 bytes of the distributed exchange have landed in this CTA's slice.  The
//          writers use st.async (asynchronous DSMEM store that reports its bytes to the receiver's mbarrier), the
//          receiver's threads spin on try_wait: no fence, no block barrier, no L1 invalidate.
//   free : (inverse only) counts one arrival per warp of the whole cluster once that warp has read its slice for
//          the last time; a CTA waits on its own copy before it stores into its peers.
// One relaxed cluster barrier at kernel start makes the initialised mbarriers visible (as in CUTLASS prologues).
__device__ __forceinline__ u32 smem_u32(const void *p) { return (u32)__cvta_generic_to_shared(p); }
template <int LOGN, int CL> struct DevCluster {
  static constexpr u32 SLICE_BYTES = (u32)(NttGeom<LOGN>::N / CL) * 8u;
  static constexpr u32 WARPS = (u32)(NttGeom<LOGN>::T / CL) / 32u;
  u32 full, free_;   // shared::cta addresses of this CTA's barriers
};
template <int LOGN, int CL> struct DevSmemView {
  u64 *local;
  u32 peer32[CL];    // shared::cluster address of every rank's slice; its `full` barrier sits SLICE_BYTES behind
  __device__ __forceinline__ void put(int r, u32 idx, u64 v) const {
    const u32 a = peer32[r] + (idx << 3), mb = peer32[r] + DevCluster<LOGN, CL>::SLICE_BYTES;
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];" ::"r"(a), "l"(v), "r"(mb) : "memory");
  }
};
template <int LOGN, int CL> __device__ __forceinline__ DevSmemView<LOGN, CL> dev_smem_view(u64 *sm) {
  DevSmemView<LOGN, CL> v;
  v.local = sm;
  const u32 base = smem_u32(sm);
#pragma unroll
  for (int r = 0; r < CL; r++) asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(v.peer32[r]) : "r"(base), "r"(r));
  return v;
}
template <int LOGN, int CL> struct DevHooks {
  DevCluster<LOGN, CL> c;
  __device__ __forceinline__ explicit DevHooks(u64 *sm) {
    c.full = smem_u32(sm) + DevCluster<LOGN, CL>::SLICE_BYTES;
    c.free_ = c.full + 8;
  }
  __device__ __forceinline__ void start() const {
    if (CL == 1) return;
    if (threadIdx.x == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(c.full) : "memory");
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(c.free_), "r"((u32)CL * DevCluster<LOGN, CL>::WARPS) : "memory");
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      // the one arrival of `full`, announcing the bytes the peers (and this CTA) will deliver
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(c.full), "r"(DevCluster<LOGN, CL>::SLICE_BYTES) : "memory");
    }
    asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
  }
  __device__ __forceinline__ void ready() const {
    if (CL > 1) asm volatile("barrier.cluster.wait.aligned;" ::: "memory");
  }
  __device__ __forceinline__ void released() const {
    if (CL == 1) return;
    __syncwarp();
    if ((threadIdx.x & 31u) == 0) {
      const u32 mine = c.free_;
#pragma unroll
      for (int r = 0; r < CL; r++) {
        u32 remote;
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(mine), "r"(r));
        asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
      }
    }
  }
  static __device__ __forceinline__ void spin(u32 bar) {
    asm volatile("{\n\t.reg .pred p;\n\tEVAB_WAIT:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@!p bra EVAB_WAIT;\n\t}" ::"r"(bar) : "memory");
  }
  __device__ __forceinline__ void acquire_free() const { if (CL > 1) spin(c.free_); }
  __device__ __forceinline__ void landed() const { if (CL > 1) spin(c.full); }
};
// barrier after a phase: 0 = block, 1 = the distributed exchange has landed in this CTA's slice, 2 = none
template <int LOGN, int CL> struct DevSync {
  DevHooks<LOGN, CL> hk;
  __device__ __forceinline__ void operator()(int kind) const {
    if (CL > 1 && kind == 1) hk.landed();
    else if (kind != 2) __syncthreads();
  }
};
template <int CL> __device__ __forceinline__ SmemView<CL> smem_view(u64 *sm) {   // CL = 1: plain view
  SmemView<CL> v;
  v.local = sm;
#pragma unroll
  for (int r = 0; r < CL; r++) v.peer[r] = sm;
  return v;
}
// CL = 1: one CTA = one residue, T = N/16 threads, 64 registers.
// CL = 2, 4, 8: one residue over a cluster of CL CTAs of T/CL threads (ntt_core.cuh), CL CTAs per SM more.
template <int LOGN, int PRO, int EPI, int CL, int AR>
__global__ void __launch_bounds__(NttGeom<LOGN>::T / CL, 1024 / (NttGeom<LOGN>::T / CL)) k_ntt_fwd(const NttLaunch L, const long long bstride) {
  extern __shared__ __align__(16) u64 sm[];
  typedef FwdBody<LOGN, PRO, EPI, CL, AR> B;
  const NttJob J = ntt_job_qr(L, blockIdx.y, blockIdx.x / CL, blockIdx.x % CL, (long long)blockIdx.z * bstride);
  if (J.skip) return;
  NttState S;
  const u32 tid = threadIdx.x;
  if (PRO == PRO_PLAIN && EPI == EPI_STORE && J.bcast) { fwd_const_poly<LOGN>(J, B::vtid(J, tid)); return; }
  const DevHooks<LOGN, CL> hk(sm);
  hk.start();   // relaxed cluster arrive; the matching wait sits right before the first store into a peer (hk.ready in phase 0)
  if constexpr (CL > 1) PhaseLoop<B, 0, B::NPH>::run(S, L, J, tid, dev_smem_view<LOGN, CL>(sm), DevSync<LOGN, CL>{hk}, hk);
  else PhaseLoop<B, 0, B::NPH>::run(S, L, J, tid, smem_view<1>(sm), DevSync<LOGN, CL>{hk}, hk);
  B::phE(S, L, J, tid);
}
template <int LOGN, int PRO, int EPI, int CL, int AR>
__global__ void __launch_bounds__(NttGeom<LOGN>::T / CL, 1024 / (NttGeom<LOGN>::T / CL)) k_ntt_inv(const NttLaunch L, const long long bstride) {
  extern __shared__ __align__(16) u64 sm[];
  typedef InvBody<LOGN, PRO, EPI, CL, AR> B;
  const NttJob J = ntt_job_qr(L, blockIdx.y, blockIdx.x / CL, blockIdx.x % CL, (long long)blockIdx.z * bstride);
  if (J.skip) return;
  NttState S;
  const DevHooks<LOGN, CL> hk(sm);
  hk.start();   // the matching wait sits right before this CTA first signals its peers (hk.ready in the pass-1 phase)
  if constexpr (CL > 1) PhaseLoop<B, 0, B::NPH>::run(S, L, J, threadIdx.x, dev_smem_view<LOGN, CL>(sm), DevSync<LOGN, CL>{hk}, hk);
  else PhaseLoop<B, 0, B::NPH>::run(S, L, J, threadIdx.x, smem_view<1>(sm), DevSync<LOGN, CL>{hk}, hk);
}
template <int OP> __global__ void __launch_bounds__(256) k_dyadic(const DyArgs A, const long long bstride) {
  const long long off = (long long)blockIdx.z * bstride;
  for (int j = (blockIdx.x * blockDim.x + threadIdx.x) * 4; j < A.N; j += gridDim.x * blockDim.x * 4)
    dyadic_elem<OP>(A, blockIdx.y, j, off);
}
__global__ void __launch_bounds__(256) k_mul_plain(const DyArgs A, const long long bstride) {
  const long long off = (long long)blockIdx.z * bstride;
  for (int j = (blockIdx.x * blockDim.x + threadIdx.x) * 4; j < A.N; j += gridDim.x * blockDim.x * 4)
    mulpt_elem(A, blockIdx.y, j, off);
}
template <bool SQ> __global__ void __launch_bounds__(256) k_mul_ct(const MulArgs A, const long long bstride) {
  const long long off = (long long)blockIdx.z * bstride;
  for (int j = (blockIdx.x * blockDim.x + threadIdx.x) * 2; j < A.N; j += gridDim.x * blockDim.x * 2)
    mulct_elem<SQ>(A, blockIdx.y, j, off);
}
__global__ void __launch_bounds__(256) k_sum_terms(const SumArgs A, const long long bstride) {
  const long long off = (long long)blockIdx.z * bstride;
  for (int j = (blockIdx.x * blockDim.x + threadIdx.x) * 2; j < A.N; j += gridDim.x * blockDim.x * 2)
    sum_terms_elem(A, blockIdx.y, j, off);
}
__global__ void __launch_bounds__(256) k_ks_inner(const IpArgs A, const long long bstride) {
  const long long off = (long long)blockIdx.z * bstride;   // parameters stay in constant memory: no local copy
  for (int j = (blockIdx.x * blockDim.x + threadIdx.x) * 2; j < A.N; j += gridDim.x * blockDim.x * 2)
    ks_inner_elem(A, blockIdx.y, j, off);
}
__global__ void __launch_bounds__(256) k_scale_c0(const u64 *c0, u64 *ext, const PrimeDev *primes, int ell, int k, int N, const long long bstride) {
  const long long off = (long long)blockIdx.z * bstride;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < N; j += gridDim.x * blockDim.x) scale_c0_elem(c0, ext, primes, ell, k, N, blockIdx.y, j, off);
}
__global__ void __launch_bounds__(256, 4) k_rot_many(const RotManyArgs A, const long long bstride) {
  const long long off = (long long)blockIdx.z * bstride;
  const int i = blockIdx.y / (A.ell + 1), mi = blockIdx.y % (A.ell + 1);
  for (int j = (blockIdx.x * blockDim.x + threadIdx.x) * 2; j < A.N; j += gridDim.x * blockDim.x * 2) rot_many_elem(A, i, mi, j, off);
}
// block = 32 element pairs x n rotations (one warp per rotation); the parts meet in shared memory, one output sum at a time
__global__ void __launch_bounds__(32 * LRS_MAX) k_lazy_rotsum(const LazyRotSumArgs A, const long long bstride) {
  __shared__ u64 sh[LRS_MAX][4][32];
  const long long off = (long long)blockIdx.z * bstride;
  const int j = (blockIdx.x * 32 + threadIdx.x) * 2, i = threadIdx.y, mi = blockIdx.y;
  u64 part[LRS_OUT][4];
  lazy_rotsum_part(A, mi, j, i, off, part);
#pragma unroll
  for (int o = 0; o < LRS_OUT; o++) {
    if (o >= A.nout) break;
    if (o) __syncthreads();
#pragma unroll
    for (int c = 0; c < 4; c++) sh[i][c][threadIdx.x] = part[o][c];
    __syncthreads();
    if (i == o % A.n) {
      u64 sum[4];
#pragma unroll
      for (int c = 0; c < 4; c++) { u64 s = 0; for (int r = 0; r < A.n; r++) s += sh[r][c][threadIdx.x]; sum[c] = s; }
      lazy_rotsum_store(A, mi, j, o, off, sum);
    }
  }
}
__global__ void __launch_bounds__(256) k_hoist_const(const HoistConstArgs A) {
  for (int j = (blockIdx.x * blockDim.x + threadIdx.x) * 2; j < A.N; j += gridDim.x * blockDim.x * 2) hoist_const_elem(A, blockIdx.y, j);
}
__global__ void __launch_bounds__(256) k_hoist_indicator(u64 *out, const u32 *ctab, int N) {
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < N; j += gridDim.x * blockDim.x) hoist_indicator_elem(out, ctab, N, blockIdx.y, j);
}
__global__ void __launch_bounds__(256) k_dec_compose(const DecArgs A) {
  for (u32 j = blockIdx.x * blockDim.x + threadIdx.x; j < A.N; j += gridDim.x * blockDim.x) dec_compose(A, j);
}
__global__ void __launch_bounds__(256) k_dec_fft(const DecArgs A, u32 m) {
  for (u32 b = blockIdx.x * blockDim.x + threadIdx.x; b < A.N / 2; b += gridDim.x * blockDim.x) dec_fft_bfly(A, m, b);
}
__global__ void __launch_bounds__(256) k_dec_gather(const DecArgs A) {
  for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < A.N / 2; i += gridDim.x * blockDim.x) dec_gather(A, i);
}
__global__ void __launch_bounds__(256) k_enc_scatter(const EncBatch B, const long long bstride, const long long vstride) {
  for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < B.N / 2; i += gridDim.x * blockDim.x)
    enc_scatter(B, blockIdx.y, i, (long long)blockIdx.z * bstride, (long long)blockIdx.z * vstride);
}
__global__ void __launch_bounds__(256) k_enc_fft(const EncBatch B, u32 g, int nstages, const long long bstride) {
  for (u32 t = blockIdx.x * blockDim.x + threadIdx.x; t < B.N / 8; t += gridDim.x * blockDim.x)
    enc_fft8(B, blockIdx.y, t, g, nstages, (long long)blockIdx.z * bstride);
}
__global__ void __launch_bounds__(256) k_enc_round(const EncBatch B, const long long bstride) {
  for (u32 j = blockIdx.x * blockDim.x + threadIdx.x; j < B.N; j += gridDim.x * blockDim.x)
    enc_round(B, blockIdx.y, j, (long long)blockIdx.z * bstride);
}
__global__ void __launch_bounds__(256) k_enc_uniform(const EncUniform B, const long long bstride) {
  const u32 e = blockIdx.y / B.ell, i = blockIdx.y % B.ell;
  for (u32 j = (blockIdx.x * blockDim.x + threadIdx.x) * 2; j < B.N; j += gridDim.x * blockDim.x * 2)
    enc_uniform_elem(B, e, i, j, (long long)blockIdx.z * bstride);
}

// ---------------------------------------------------------------------------
// CUDA backend
// ---------------------------------------------------------------------------
// CTAs per residue (evab_set_ntt_cluster); 0 = automatic: 128-thread CTAs, i.e. N/2048 CTAs per residue
// (N=16384: 8, 8192: 4, 4096: 2).  Sobel ops/s and single-instance latency at N=16384:
// 1: 414k 1.11 ms, 2: 444k 0.76, 4: 477k 0.57, 8: 494k 0.49; N=8192 is fastest at 4 (64-thread CTAs lose).
static std::atomic<int> g_ntt_cluster{0};   // default for contexts created afterwards (evab_set_ntt_cluster)
template <int LOGN> static int ntt_cluster_for(int requested) {
  int most = NttGeom<LOGN>::T / 128 < 1 ? 1 : NttGeom<LOGN>::T / 128;   // keep CTAs at >= 128 threads
  if (most > 8) most = 8;                                                 // portable cluster size
  int cl = requested ? requested : most;
  if (cl > most) cl = most;
  if (LOGN == 15 && cl < 2) cl = 2;                                       // 2048 threads never fit one CTA
  return cl;
}

// grid = (inner * cluster, q, batch instance); more than 65535 q's are issued in slices
template <class K> static int launch_ntt(K kernel, const NttLaunch &L0, size_t jobs, int threads, size_t smem, int cluster, cudaStream_t st, std::atomic<bool> *done) {
  int dev = 0;
  cudaGetDevice(&dev);
  if (!done[dev & 63].load()) {
    CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    done[dev & 63].store(true);
  }
  const size_t inner = (size_t)L0.inner, nq = (jobs + inner - 1) / inner;
  if (nq * inner != jobs) return fail("NTT launch: job count must be a multiple of the inner dimension");
  if (nq > 65535 && L0.prime_on_q) return fail("NTT launch: too many polynomials for a prime-per-polynomial launch");
  for (size_t q0 = 0; q0 < nq; q0 += 65535) {
    NttLaunch L = L0;
    const long long o = (long long)q0;
    L.src += o * L.src_sq; L.dst += o * L.dst_sq;
    if (L.aux0) L.aux0 += o * L.aux0_sq;
    if (L.aux1) L.aux1 += o * L.aux1_sq;
    if (L.cflags) L.cflags += o;
    L.aux1_polys = L0.aux1_polys - (int)(o > (1 << 30) ? (1 << 30) : o);
    const size_t n = nq - q0 < 65535 ? nq - q0 : 65535;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(inner * cluster), (unsigned)n, (unsigned)g_batch.batch);
    cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = cluster; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    CUDA_OK(cudaLaunchKernelEx(&cfg, kernel, L, g_batch.stride));
  }
  return 0;
}
template <int LOGN, int PRO, int EPI, int CL, int AR> static int launch_fwd_m(const NttLaunch &L, size_t jobs, cudaStream_t st) {
  static std::atomic<bool> done[64];
  return launch_ntt(k_ntt_fwd<LOGN, PRO, EPI, CL, AR>, L, jobs, NttGeom<LOGN>::T / CL, (size_t)NttGeom<LOGN>::N / CL * sizeof(u64) + (CL > 1 ? 16 : 0), CL, st, done);
}
template <int LOGN, int CL, int AR> static int launch_fwd_c(const NttLaunch &L, size_t jobs, cudaStream_t st) {
  if (L.pro == PRO_PLAIN && L.epi == EPI_STORE) return launch_fwd_m<LOGN, PRO_PLAIN, EPI_STORE, CL, AR>(L, jobs, st);
  if (L.pro == PRO_MODRED && L.epi == EPI_STORE) return launch_fwd_m<LOGN, PRO_MODRED, EPI_STORE, CL, AR>(L, jobs, st);
  if (L.pro == PRO_MODRED && L.epi == EPI_STORE_LAZY) return launch_fwd_m<LOGN, PRO_MODRED, EPI_STORE_LAZY, CL, AR>(L, jobs, st);
  if (L.pro == PRO_MODRED && L.epi == EPI_DIVROUND) return launch_fwd_m<LOGN, PRO_MODRED, EPI_DIVROUND, CL, AR>(L, jobs, st);
  if (L.pro == PRO_MODRED_SG && L.epi == EPI_STORE_LAZY) return launch_fwd_m<LOGN, PRO_MODRED_SG, EPI_STORE_LAZY, CL, AR>(L, jobs, st);
  return fail("unsupported forward NTT prologue/epilogue combination");
}
template <int LOGN, int AR> static int launch_fwd_a(const NttLaunch &L, size_t jobs, cudaStream_t st, int cluster) {
  if constexpr (LOGN >= 12) {
    const int cl = ntt_cluster_for<LOGN>(cluster);
    if constexpr (LOGN >= 13) {
      if (cl == 4) return launch_fwd_c<LOGN, 4, AR>(L, jobs, st);
      if (cl == 8) return launch_fwd_c<LOGN, 8, AR>(L, jobs, st);
    }
    if (cl == 2 || LOGN == 15) return launch_fwd_c<LOGN, 2, AR>(L, jobs, st);
  }
  if constexpr (LOGN <= 14) return launch_fwd_c<LOGN, 1, AR>(L, jobs, st);
  return fail("unsupported cluster size");
}
template <int LOGN> static int launch_fwd_t(const NttLaunch &L, size_t jobs, cudaStream_t st, bool fold, int cluster) {
  return fold ? launch_fwd_a<LOGN, 1>(L, jobs, st, cluster) : launch_fwd_a<LOGN, 0>(L, jobs, st, cluster);
}
template <int LOGN, int PRO, int EPI, int CL, int AR> static int launch_inv_m(const NttLaunch &L, size_t jobs, cudaStream_t st) {
  static std::atomic<bool> done[64];
  return launch_ntt(k_ntt_inv<LOGN, PRO, EPI, CL, AR>, L, jobs, NttGeom<LOGN>::T / CL, (size_t)NttGeom<LOGN>::N / CL * sizeof(u64) + (CL > 1 ? 16 : 0), CL, st, done);
}
template <int LOGN, int CL, int AR> static int launch_inv_c(const NttLaunch &L, size_t jobs, cudaStream_t st) {
  if (L.pro == PRO_PLAIN && L.epi == EPI_STORE) return launch_inv_m<LOGN, PRO_PLAIN, EPI_STORE, CL, AR>(L, jobs, st);
  if (L.pro == PRO_PLAIN && L.epi == EPI_ADDHALF) return launch_inv_m<LOGN, PRO_PLAIN, EPI_ADDHALF, CL, AR>(L, jobs, st);
  if (L.pro == PRO_GATHER && L.epi == EPI_STORE) return launch_inv_m<LOGN, PRO_GATHER, EPI_STORE, CL, AR>(L, jobs, st);
  if (L.pro == PRO_PLAIN && L.epi == EPI_STORE_ZFLAG) return launch_inv_m<LOGN, PRO_PLAIN, EPI_STORE_ZFLAG, CL, AR>(L, jobs, st);
  return fail("unsupported inverse NTT prologue/epilogue combination");
}
template <int LOGN, int AR> static int launch_inv_a(const NttLaunch &L, size_t jobs, cudaStream_t st, int cluster) {
  if constexpr (LOGN >= 12) {
    const int cl = ntt_cluster_for<LOGN>(cluster);
    if constexpr (LOGN >= 13) {
      if (cl == 4) return launch_inv_c<LOGN, 4, AR>(L, jobs, st);
      if (cl == 8) return launch_inv_c<LOGN, 8, AR>(L, jobs, st);
    }
    if (cl == 2 || LOGN == 15) return launch_inv_c<LOGN, 2, AR>(L, jobs, st);
  }
  if constexpr (LOGN <= 14) return launch_inv_c<LOGN, 1, AR>(L, jobs, st);
  return fail("unsupported cluster size");
}
template <int LOGN> static int launch_inv_t(const NttLaunch &L, size_t jobs, cudaStream_t st, bool fold, int cluster) {
  return fold ? launch_inv_a<LOGN, 1>(L, jobs, st, cluster) : launch_inv_a<LOGN, 0>(L, jobs, st, cluster);
}

struct CudaBE {
  const evab_ctx *c;
  cudaStream_t st;
  int error(const char *m) { return fail(m); }
  void count(unsigned n = 1) const { c->launches.fetch_add(n, std::memory_order_relaxed); }
  dim3 grid(int rows, int per_thread = 2) const {
    int per_row = (int)(c->v.N / per_thread / 256);
    if (per_row < 1) per_row = 1;
    return dim3(per_row, rows, g_batch.batch);
  }
  int fwd(const NttLaunch &L, size_t jobs) {
    count();
    const bool fold = c->v.arith == 0 && ntt_launch_folds(L, jobs, c->v.foldmask);
    switch (c->v.logN) {
      case 10: return launch_fwd_t<10>(L, jobs, st, fold, c->ntt_cluster);
      case 11: return launch_fwd_t<11>(L, jobs, st, fold, c->ntt_cluster);
      case 12: return launch_fwd_t<12>(L, jobs, st, fold, c->ntt_cluster);
      case 13: return launch_fwd_t<13>(L, jobs, st, fold, c->ntt_cluster);
      case 14: return launch_fwd_t<14>(L, jobs, st, fold, c->ntt_cluster);
      case 15: return launch_fwd_t<15>(L, jobs, st, fold, c->ntt_cluster);
    }
    return fail("unsupported N");
  }
  int inv(const NttLaunch &L, size_t jobs) {
    count();
    const bool fold = c->v.arith == 0 && ntt_launch_folds(L, jobs, c->v.foldmask);
    switch (c->v.logN) {
      case 10: return launch_inv_t<10>(L, jobs, st, fold, c->ntt_cluster);
      case 11: return launch_inv_t<11>(L, jobs, st, fold, c->ntt_cluster);
      case 12: return launch_inv_t<12>(L, jobs, st, fold, c->ntt_cluster);
      case 13: return launch_inv_t<13>(L, jobs, st, fold, c->ntt_cluster);
      case 14: return launch_inv_t<14>(L, jobs, st, fold, c->ntt_cluster);
      case 15: return launch_inv_t<15>(L, jobs, st, fold, c->ntt_cluster);
    }
    return fail("unsupported N");
  }
  int dyadic(int op, const DyArgs &A) {
    count();
    dim3 g = grid(A.sout * A.ell, 4);
    switch (op) {
      case DY_ADD: k_dyadic<DY_ADD><<<g, 256, 0, st>>>(A, g_batch.stride); break;
      case DY_SUB: k_dyadic<DY_SUB><<<g, 256, 0, st>>>(A, g_batch.stride); break;
      case DY_NEG: k_dyadic<DY_NEG><<<g, 256, 0, st>>>(A, g_batch.stride); break;
      case DY_COPY: k_dyadic<DY_COPY><<<g, 256, 0, st>>>(A, g_batch.stride); break;
      default: k_mul_plain<<<grid(A.ell, 4), 256, 0, st>>>(A, g_batch.stride); break;   // one thread: the same coefficients of every polynomial
    }
    CUDA_OK(cudaGetLastError());
    return 0;
  }
  int sum(const SumArgs &A) {
    count();
    k_sum_terms<<<grid(A.sout * A.ell), 256, 0, st>>>(A, g_batch.stride);
    CUDA_OK(cudaGetLastError());
    return 0;
  }
  int mulct(bool sq, const MulArgs &A) {
    count();
    if (sq) k_mul_ct<true><<<grid(A.ell), 256, 0, st>>>(A, g_batch.stride);
    else k_mul_ct<false><<<grid(A.ell), 256, 0, st>>>(A, g_batch.stride);
    CUDA_OK(cudaGetLastError());
    return 0;
  }
  int inner(const IpArgs &A) {
    count();
    k_ks_inner<<<grid(A.ell + 1), 256, 0, st>>>(A, g_batch.stride);
    CUDA_OK(cudaGetLastError());
    return 0;
  }
  int scale_c0(const u64 *c0, u64 *ext, int ell) {
    count();
    k_scale_c0<<<grid(ell, 1), 256, 0, st>>>(c0, ext, c->v.primes, ell, c->v.k, (int)c->v.N, g_batch.stride);
    CUDA_OK(cudaGetLastError());
    return 0;
  }
  int rot_many(const RotManyArgs &A) {
    count();
    k_rot_many<<<grid(A.n * (A.ell + 1)), 256, 0, st>>>(A, g_batch.stride);
    CUDA_OK(cudaGetLastError());
    return 0;
  }
  int lazy_rotsum(const LazyRotSumArgs &A) {
    count();
    k_lazy_rotsum<<<dim3(A.N / 64, A.ell + 1, g_batch.batch), dim3(32, A.n), 0, st>>>(A, g_batch.stride);
    CUDA_OK(cudaGetLastError());
    return 0;
  }
  int hoist_indicator(u64 *out, const u32 *ctab, int N, int rows) {
    count();
    k_hoist_indicator<<<dim3((N + 255) / 256, rows), 256, 0, st>>>(out, ctab, N);
    CUDA_OK(cudaGetLastError());
    return 0;
  }
  int hoist_const(const HoistConstArgs &A) {
    count();
    k_hoist_const<<<dim3((A.N / 2 + 255) / 256, A.ell + 1), 256, 0, st>>>(A);
    CUDA_OK(cudaGetLastError());
    return 0;
  }
  int dec_compose(const DecArgs &A) { count(); k_dec_compose<<<(A.N + 255) / 256, 256, 0, st>>>(A); CUDA_OK(cudaGetLastError()); return 0; }
  int dec_fft(const DecArgs &A, u32 m) { count(); k_dec_fft<<<(A.N / 2 + 255) / 256, 256, 0, st>>>(A, m); CUDA_OK(cudaGetLastError()); return 0; }
  int dec_gather(const DecArgs &A) { count(); k_dec_gather<<<(A.N / 2 + 255) / 256, 256, 0, st>>>(A); CUDA_OK(cudaGetLastError()); return 0; }
  int enc_scatter(const EncBatch &B) {
    count();
    k_enc_scatter<<<dim3((B.N / 2 + 255) / 256, B.count, g_batch.batch), 256, 0, st>>>(B, g_batch.stride, g_batch.vstride);
    CUDA_OK(cudaGetLastError());
    return 0;
  }
  int enc_fft(const EncBatch &B, u32 g, int ns) {
    count();
    k_enc_fft<<<dim3((B.N / 8 + 255) / 256, B.count, g_batch.batch), 256, 0, st>>>(B, g, ns, g_batch.stride);
    CUDA_OK(cudaGetLastError());
    return 0;
  }
  int enc_uniform(const EncUniform &B) {
    count();
    k_enc_uniform<<<dim3((B.N / 2 + 255) / 256, B.count * B.ell, g_batch.batch), 256, 0, st>>>(B, g_batch.stride);
    CUDA_OK(cudaGetLastError());
    return 0;
  }
  int enc_round(const EncBatch &B) {
    count();
    k_enc_round<<<dim3((B.N + 255) / 256, B.count, g_batch.batch), 256, 0, st>>>(B, g_batch.stride);
    CUDA_OK(cudaGetLastError());
    return 0;
  }
};

// ---------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------
static void ctx_free(evab_ctx *c) {
  for (auto &kv : c->perms) cudaFree(kv.second);
  for (auto &kv : c->cperms) cudaFree(kv.second);
  cudaFree(c->d_tw); cudaFree(c->d_primes); cudaFree(c->d_qinv); cudaFree(c->d_qinv_f); cudaFree(c->d_halfmod); cudaFree(c->d_zeros);
  cudaFree(c->d_roots); cudaFree(c->d_slot); cudaFree(c->d_pow2);
  delete c;
}
// device copy of a host table; on failure the caller frees the whole context (no leak on any error path)
template <class T> static int to_device(T **dst, const T *src, size_t n) {
  CUDA_OK(cudaMalloc(dst, n * sizeof(T)));
  CUDA_OK(cudaMemcpy(*dst, src, n * sizeof(T), cudaMemcpyHostToDevice));
  return 0;
}
static int ctx_build(evab_ctx *c, uint64_t N, int logN, const uint64_t *primes, int k) {
  const size_t tw_elems = (size_t)k * 4 * N;   // Shoup rows + fold rows (host_tables.hpp)
  CUDA_OK(cudaMalloc(&c->d_tw, tw_elems * sizeof(u64x2)));
  evab_host::Tables T;
  const char *err = evab_host::build_tables(N, logN, primes, k, c->d_tw, T);
  if (err[0]) return fail(std::string("evab_ctx_create: ") + err);
  CUDA_OK(cudaMemcpy(c->d_tw, T.tw.data(), tw_elems * sizeof(u64x2), cudaMemcpyHostToDevice));
  if (to_device(&c->d_primes, T.pd.data(), (size_t)k)) return 1;
  if (to_device(&c->d_qinv, T.qinv.data(), T.qinv.size())) return 1;
  if (to_device(&c->d_qinv_f, T.qinv_f.data(), T.qinv_f.size())) return 1;
  if (to_device(&c->d_halfmod, T.halfmod.data(), T.halfmod.size())) return 1;
  CUDA_OK(cudaMalloc(&c->d_zeros, 32 * sizeof(u64)));
  CUDA_OK(cudaMemset(c->d_zeros, 0, 32 * sizeof(u64)));
  if (to_device(&c->d_roots, reinterpret_cast<const cplx *>(T.roots.data()), (size_t)N)) return 1;
  if (to_device(&c->d_slot, T.slot_index.data(), (size_t)N)) return 1;
  if (to_device(&c->d_pow2, T.pow2.data(), T.pow2.size())) return 1;
  c->v.roots = c->d_roots; c->v.slot_index = c->d_slot; c->v.pow2 = c->d_pow2;
  c->v.N = N; c->v.logN = logN; c->v.k = k;
  c->v.primes = c->d_primes; c->v.qinv = c->d_qinv; c->v.qinv_f = c->d_qinv_f; c->v.halfmod = c->d_halfmod; c->v.zeros = c->d_zeros;
  c->v.foldmask = T.foldmask;
  c->fold_host = T.fp; c->v.fold_host = c->fold_host.data();
  c->v.arith = getenv("EVAB_NTT_SHOUP") ? 1 : 0;
  return 0;
}
extern "C" int evab_ctx_create(uint64_t N, const uint64_t *primes, int k, int device, evab_ctx **out) {
  if (!out) return fail("evab_ctx_create: null out");
  *out = nullptr;
  int logN = 0;
  while ((1ull << logN) < N) logN++;
  if ((1ull << logN) != N || logN < 10 || logN > 15) return fail("evab_ctx_create: N must be a power of two in [2^10, 2^15]");
  if (k < 1 || k > 24) return fail("evab_ctx_create: prime count must be in [1,24]");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail("evab_ctx_create: no CUDA device (this backend has no CPU fallback)");
  if (device < 0 || device >= ndev) return fail("evab_ctx_create: bad device index");
  CUDA_OK(cudaSetDevice(device));
  cudaDeviceProp prop;
  CUDA_OK(cudaGetDeviceProperties(&prop, device));
  cudaMemPool_t pool;
  CUDA_OK(cudaDeviceGetDefaultMemPool(&pool, device));
  unsigned long long thresh = ~0ull;  // keep freed blocks cached in the pool
  cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thresh);

  evab_ctx *c = new evab_ctx();
  c->device = device; c->sms = prop.multiProcessorCount;
  c->ntt_cluster = g_ntt_cluster.load();
  c->primes.assign(primes, primes + k);
  if (ctx_build(c, N, logN, primes, k)) { ctx_free(c); return 1; }   // g_err is set; every partial allocation is released
  *out = c;
  return 0;
}
extern "C" void evab_ctx_destroy(evab_ctx *c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaDeviceSynchronize();
  ctx_free(c);
}
// per-context tuning knobs (no process-global state): CTAs per residue, NTT arithmetic
extern "C" int evab_ctx_set_ntt_cluster(evab_ctx *c, int cl) {
  if (cl != 0 && cl != 1 && cl != 2 && cl != 4 && cl != 8) return fail("evab_ctx_set_ntt_cluster: 0 (automatic), 1, 2, 4 or 8 CTAs per residue");
  c->ntt_cluster = cl;
  return 0;
}
extern "C" int evab_ctx_set_ntt_arith(evab_ctx *c, int mode) {
  if (mode != 0 && mode != 1) return fail("evab_ctx_set_ntt_arith: 0 (two-row fold where the primes allow it) or 1 (lazy Shoup everywhere)");
  c->v.arith = mode;
  return 0;
}
extern "C" unsigned evab_ctx_foldmask(const evab_ctx *c) { return c->v.foldmask; }
extern "C" uint64_t evab_ctx_N(const evab_ctx *c) { return c->v.N; }
extern "C" int evab_ctx_k(const evab_ctx *c) { return c->v.k; }
extern "C" int evab_ctx_device(const evab_ctx *c) { return c->device; }
extern "C" int evab_ctx_sm_count(const evab_ctx *c) { return c->sms; }
extern "C" int evab_mem_info(evab_ctx *c, size_t *free_bytes, size_t *total_bytes) {
  CUDA_OK(cudaSetDevice(c->device));
  CUDA_OK(cudaMemGetInfo(free_bytes, total_bytes));
  return 0;
}

extern "C" int evab_malloc(evab_ctx *c, size_t bytes, void **p, void *stream) {
  CUDA_OK(cudaSetDevice(c->device));
  CUDA_OK(cudaMallocAsync(p, bytes ? bytes : 8, S(stream)));
  return 0;
}
extern "C" int evab_free(evab_ctx *c, void *p, void *stream) {
  CUDA_OK(cudaSetDevice(c->device));
  CUDA_OK(cudaFreeAsync(p, S(stream)));
  return 0;
}
extern "C" int evab_upload(evab_ctx *c, void *d, const void *h, size_t bytes, void *stream) {
  CUDA_OK(cudaSetDevice(c->device));
  CUDA_OK(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, S(stream)));
  return 0;
}
extern "C" int evab_download(evab_ctx *c, void *h, const void *d, size_t bytes, void *stream) {
  CUDA_OK(cudaSetDevice(c->device));
  CUDA_OK(cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, S(stream)));
  return 0;
}
extern "C" int evab_host_alloc(size_t bytes, void **p) {
  CUDA_OK(cudaHostAlloc(p, bytes ? bytes : 8, cudaHostAllocPortable));
  return 0;
}
extern "C" int evab_host_free(void *p) {
  CUDA_OK(cudaFreeHost(p));
  return 0;
}
extern "C" int evab_sync(evab_ctx *c, void *stream) {
  CUDA_OK(cudaSetDevice(c->device));
  CUDA_OK(cudaStreamSynchronize(S(stream)));
  return 0;
}
extern "C" int evab_stream_create(evab_ctx *c, void **stream) {
  CUDA_OK(cudaSetDevice(c->device));
  cudaStream_t s;
  CUDA_OK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
  *stream = (void *)s;
  return 0;
}
extern "C" int evab_stream_destroy(evab_ctx *c, void *stream) {
  CUDA_OK(cudaSetDevice(c->device));
  CUDA_OK(cudaStreamDestroy(S(stream)));
  return 0;
}
extern "C" int evab_event_create(evab_ctx *c, void **event) {
  CUDA_OK(cudaSetDevice(c->device));
  cudaEvent_t e;
  CUDA_OK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  *event = (void *)e;
  return 0;
}
extern "C" int evab_event_destroy(evab_ctx *c, void *event) {
  CUDA_OK(cudaSetDevice(c->device));
  CUDA_OK(cudaEventDestroy((cudaEvent_t)event));
  return 0;
}
extern "C" int evab_event_record(evab_ctx *c, void *event, void *stream) {
  CUDA_OK(cudaEventRecord((cudaEvent_t)event, S(stream)));
  return 0;
}
extern "C" int evab_stream_wait_event(evab_ctx *c, void *stream, void *event) {
  CUDA_OK(cudaStreamWaitEvent(S(stream), (cudaEvent_t)event, 0));
  return 0;
}
extern "C" int evab_graph_begin(evab_ctx *c, void *stream) {
  CUDA_OK(cudaSetDevice(c->device));
  CUDA_OK(cudaStreamBeginCapture(S(stream), cudaStreamCaptureModeThreadLocal));
  return 0;
}
extern "C" int evab_graph_end(evab_ctx *c, void *stream, void **graph_exec) {
  cudaGraph_t g = nullptr;
  CUDA_OK(cudaStreamEndCapture(S(stream), &g));
  cudaGraphExec_t ge = nullptr;
  cudaError_t e = cudaGraphInstantiate(&ge, g, 0);
  cudaGraphDestroy(g);
  if (e != cudaSuccess) return fail(std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e));
  *graph_exec = (void *)ge;
  return 0;
}
extern "C" int evab_graph_launch(evab_ctx *c, void *graph_exec, void *stream) {
  CUDA_OK(cudaGraphLaunch((cudaGraphExec_t)graph_exec, S(stream)));
  return 0;
}
extern "C" int evab_graph_destroy(evab_ctx *c, void *graph_exec) {
  CUDA_OK(cudaGraphExecDestroy((cudaGraphExec_t)graph_exec));
  return 0;
}
extern "C" uint64_t evab_launch_count(const evab_ctx *c) { return c->launches.load(); }

// ---------------------------------------------------------------------------
// ops
// ---------------------------------------------------------------------------
#define BE_BEGIN                         \
  CUDA_OK(cudaSetDevice(c->device));     \
  CudaBE be{c, S(stream)};

extern "C" int evab_ntt_fwd(evab_ctx *c, uint64_t *d, size_t count, const int *pidx, int np, void *stream) {
  BE_BEGIN return ntt_batch_impl(be, c->v, false, d, count, pidx, np);
}
extern "C" int evab_ntt_inv(evab_ctx *c, uint64_t *d, size_t count, const int *pidx, int np, void *stream) {
  BE_BEGIN return ntt_batch_impl(be, c->v, true, d, count, pidx, np);
}
extern "C" int evab_set_ntt_cluster(int cl) {
  if (cl != 0 && cl != 1 && cl != 2 && cl != 4 && cl != 8) return fail("evab_set_ntt_cluster: 0 (automatic), 1, 2, 4 or 8 CTAs per residue");
  g_ntt_cluster.store(cl);   // default of contexts created from now on; existing contexts keep theirs
  return 0;
}
extern "C" int evab_encode_uniform(evab_ctx *c, int count, const double *values, const double *scales, int ell, uint64_t *out, void *stream) {
  BE_BEGIN return encode_uniform_impl(be, c->v, count, values, scales, ell, out);
}
extern "C" int evab_encode_uniform_ext(evab_ctx *c, int count, const double *values, const double *scales, int ell, int with_p, uint64_t *out, void *stream) {
  BE_BEGIN return encode_uniform_impl(be, c->v, count, values, scales, ell, out, with_p);
}
extern "C" int evab_encode_ext(evab_ctx *c, int count, const double *const *vals, const uint32_t *vec, const double *scales, int ell, int with_p,
                               uint64_t *out, void *work, void *stream) {
  BE_BEGIN return encode_impl(be, c->v, count, vals, vec, scales, ell, out, (cplx *)work, with_p);
}
extern "C" size_t evab_encode_work_bytes(const evab_ctx *c, int count) { return encode_work_bytes(c->v, count); }
extern "C" int evab_encode(evab_ctx *c, int count, const double *const *vals, const uint32_t *vec, const double *scales, int ell,
                           uint64_t *out, void *work, void *stream) {
  BE_BEGIN return encode_impl(be, c->v, count, vals, vec, scales, ell, out, (cplx *)work);
}
extern "C" size_t evab_decode_work_bytes(const evab_ctx *c, int ell) { return decode_tmp_elems(c->v, ell) * sizeof(u64) + (size_t)c->v.N * sizeof(cplx); }
extern "C" int evab_decode(evab_ctx *c, int ell, const uint64_t *pt, double scale, double *d_out, void *work, void *stream) {
  u64 *tmp = (u64 *)work;
  cplx *cw = (cplx *)(tmp + decode_tmp_elems(c->v, ell));
  BE_BEGIN return decode_impl(be, c->v, ell, c->primes.data(), pt, scale, d_out, tmp, cw);
}
extern "C" int evab_add(evab_ctx *c, int ell, uint64_t *o, const uint64_t *a, int sa, const uint64_t *b, int sb, void *stream) {
  BE_BEGIN return dyadic_impl<DY_ADD>(be, c->v, ell, o, a, sa, b, sb, 0);
}
extern "C" int evab_sub(evab_ctx *c, int ell, uint64_t *o, const uint64_t *a, int sa, const uint64_t *b, int sb, void *stream) {
  BE_BEGIN return dyadic_impl<DY_SUB>(be, c->v, ell, o, a, sa, b, sb, 0);
}
extern "C" int evab_add_plain(evab_ctx *c, int ell, uint64_t *o, const uint64_t *a, int sa, const uint64_t *pt, void *stream) {
  BE_BEGIN return dyadic_impl<DY_ADD>(be, c->v, ell, o, a, sa, pt, 1, 1);
}
extern "C" int evab_sub_plain(evab_ctx *c, int ell, uint64_t *o, const uint64_t *a, int sa, const uint64_t *pt, void *stream) {
  BE_BEGIN return dyadic_impl<DY_SUB>(be, c->v, ell, o, a, sa, pt, 1, 1);
}
extern "C" int evab_negate(evab_ctx *c, int ell, uint64_t *o, const uint64_t *a, int sa, void *stream) {
  BE_BEGIN return dyadic_impl<DY_NEG>(be, c->v, ell, o, a, sa, (const u64 *)nullptr, 0, 0);
}
extern "C" int evab_mul_plain(evab_ctx *c, int ell, uint64_t *o, const uint64_t *a, int sa, const uint64_t *pt, void *stream) {
  BE_BEGIN return dyadic_impl<DY_MULPT>(be, c->v, ell, o, a, sa, pt, 1, 1);
}
extern "C" int evab_sum_terms(evab_ctx *c, int ell, uint64_t *o, int n, const uint64_t *const *cts, const int *sizes, const uint64_t *const *pts, void *stream) {
  BE_BEGIN return sum_terms_impl(be, c->v, ell, o, n, cts, sizes, pts);
}
extern "C" int evab_sum_products(evab_ctx *c, int ell, uint64_t *o, int n, const uint64_t *const *cts, const int *sizes, const uint64_t *const *seconds, const int *kinds,
                                 void *stream) {
  BE_BEGIN return sum_terms_impl(be, c->v, ell, o, n, cts, sizes, seconds, kinds);
}
extern "C" int evab_mul(evab_ctx *c, int ell, uint64_t *o, const uint64_t *a, const uint64_t *b, void *stream) {
  BE_BEGIN return mulct_impl(be, c->v, false, ell, o, a, b);
}
extern "C" int evab_square(evab_ctx *c, int ell, uint64_t *o, const uint64_t *a, void *stream) {
  BE_BEGIN return mulct_impl(be, c->v, true, ell, o, a, (const u64 *)nullptr);
}
extern "C" int evab_mod_switch(evab_ctx *c, int ell, uint64_t *o, const uint64_t *a, int sa, void *stream) {
  if (ell < 2 || ell > c->v.k) return fail("mod_switch needs 2 <= ell <= k");
  BE_BEGIN return copy_impl(be, c->v, ell, ell - 1, o, a, sa);
}
extern "C" int evab_copy(evab_ctx *c, int ell, uint64_t *o, const uint64_t *a, int sa, void *stream) {
  BE_BEGIN return copy_impl(be, c->v, ell, ell, o, a, sa);
}
extern "C" size_t evab_rescale_work_bytes(const evab_ctx *c, int sa) { return rescale_work_elems(c->v, sa) * sizeof(u64); }
extern "C" int evab_rescale(evab_ctx *c, int ell, uint64_t *o, const uint64_t *a, int sa, void *work, void *stream) {
  BE_BEGIN return rescale_impl(be, c->v, ell, o, a, sa, (u64 *)work);
}
extern "C" size_t evab_keyswitch_work_bytes(const evab_ctx *c, int ell) { return keyswitch_work_elems(c->v, ell) * sizeof(u64); }
extern "C" int evab_relinearize(evab_ctx *c, int ell, uint64_t *o, const uint64_t *a, const uint64_t *key, void *work, void *stream) {
  BE_BEGIN return relinearize_impl(be, c->v, ell, o, a, key, (u64 *)work);
}
extern "C" uint64_t evab_galois_elt_from_step(uint64_t N, int steps) { return evab_host::galois_elt_from_step(N, steps); }
extern "C" int evab_galois_prepare(evab_ctx *c, uint64_t elt) {
  if (!(elt & 1) || elt >= 2 * c->v.N) return fail("galois element must be odd and < 2N");
  std::lock_guard<std::mutex> g(c->mu);
  if (c->perms.count(elt)) return 0;
  CUDA_OK(cudaSetDevice(c->device));
  std::vector<u32> tab;
  evab_host::galois_table(c->v.N, c->v.logN, elt, tab);
  u32 *d = nullptr;
  CUDA_OK(cudaMalloc(&d, tab.size() * sizeof(u32)));
  CUDA_OK(cudaMemcpy(d, tab.data(), tab.size() * sizeof(u32), cudaMemcpyHostToDevice));
  c->perms[elt] = d;
  evab_host::galois_coeff_table(c->v.N, elt, tab);
  u32 *dc = nullptr;
  CUDA_OK(cudaMalloc(&dc, tab.size() * sizeof(u32)));
  CUDA_OK(cudaMemcpy(dc, tab.data(), tab.size() * sizeof(u32), cudaMemcpyHostToDevice));
  c->cperms[elt] = dc;
  return 0;
}
extern "C" int evab_rotate_prepare(evab_ctx *c, int ell, uint64_t *hoist, const uint64_t *a, void *stream) {
  BE_BEGIN return rotate_prepare_impl(be, c->v, ell, hoist, a);
}
extern "C" int evab_rotate_prepared(evab_ctx *c, int ell, uint64_t *o, const uint64_t *a, const uint64_t *hoist, uint64_t elt, const uint64_t *key, void *work,
                                    void *stream) {
  u32 *perm = nullptr, *ctab = nullptr;
  {
    std::lock_guard<std::mutex> g(c->mu);
    auto it = c->perms.find(elt);
    if (it == c->perms.end()) return fail("evab_rotate_prepared: call evab_galois_prepare(elt) first");
    perm = it->second; ctab = c->cperms.at(elt);
  }
  BE_BEGIN return rotate_prepared_impl(be, c->v, ell, o, a, hoist, perm, ctab, key, (u64 *)work);
}
static int galois_tables(evab_ctx *c, uint64_t elt, u32 **perm, u32 **ctab, const char *who) {
  std::lock_guard<std::mutex> g(c->mu);
  auto it = c->perms.find(elt);
  if (it == c->perms.end()) return fail(std::string(who) + ": call evab_galois_prepare(elt) first");
  *perm = it->second; *ctab = c->cperms.at(elt);
  return 0;
}
extern "C" size_t evab_rotate_modup_work_bytes(const evab_ctx *c, int ell) { return rotate_modup_work_elems(c->v, ell) * sizeof(u64); }
extern "C" size_t evab_rotate_modup_ext_bytes(const evab_ctx *c, int ell) { return (size_t)(ell + 1) * ell * c->v.N * sizeof(u64); }
extern "C" size_t evab_hoist_const_bytes(const evab_ctx *c, int ell) { return hoist_const_elems(c->v, ell) * sizeof(u64); }
extern "C" int evab_rotate_modup_prepare(evab_ctx *c, int ell, uint64_t *that, uint64_t *ext, const uint64_t *a, uint64_t *zflag, void *stream) {
  BE_BEGIN return rotate_modup_prepare_impl(be, c->v, ell, that, ext, a, zflag);
}
extern "C" int evab_rotate_hoist_const(evab_ctx *c, int ell, uint64_t elt, const uint64_t *key, uint64_t *out, uint64_t *tmp, void *stream) {
  u32 *perm = nullptr, *ctab = nullptr;
  if (galois_tables(c, elt, &perm, &ctab, "evab_rotate_hoist_const")) return 1;
  BE_BEGIN return hoist_const_impl(be, c->v, ell, ctab, key, out, tmp);
}
extern "C" int evab_rotate_modup_prepared(evab_ctx *c, int ell, uint64_t *o, const uint64_t *a, const uint64_t *ext, uint64_t elt, const uint64_t *key,
                                          const uint64_t *cadd, void *work, void *stream) {
  u32 *perm = nullptr, *ctab = nullptr;
  if (galois_tables(c, elt, &perm, &ctab, "evab_rotate_modup_prepared")) return 1;
  BE_BEGIN return rotate_modup_prepared_impl(be, c->v, ell, o, a, ext, perm, key, cadd, (u64 *)work);
}
extern "C" int evab_rotate_modup_scale_c0(evab_ctx *c, int ell, uint64_t *ext, const uint64_t *a, void *stream) {
  BE_BEGIN return rotate_modup_scale_c0_impl(be, c->v, ell, ext, a);
}
extern "C" size_t evab_rotate_modup_many_work_bytes(const evab_ctx *c, int ell, int n) { return rotate_modup_many_work_elems(c->v, ell, n) * sizeof(u64); }
extern "C" int evab_rotate_modup_many(evab_ctx *c, int ell, int n, uint64_t *o, const uint64_t *a, const uint64_t *ext, const uint64_t *elts,
                                      const uint64_t *const *keys, const uint64_t *const *cadds, void *work, void *stream) {
  if (n < 1 || n > ROTMANY_MAX) return fail("evab_rotate_modup_many: 1..16 rotations");
  const u32 *perms[ROTMANY_MAX];
  for (int i = 0; i < n; i++) {
    u32 *perm = nullptr, *ctab = nullptr;
    if (galois_tables(c, elts[i], &perm, &ctab, "evab_rotate_modup_many")) return 1;
    perms[i] = perm;
  }
  BE_BEGIN return rotate_modup_many_impl(be, c->v, ell, n, o, a, ext, perms, keys, cadds, (u64 *)work);
}
extern "C" size_t evab_lazy_rotsum_work_bytes(const evab_ctx *c, int ell, int nout) { return lazy_rotsum_work_elems(c->v, ell, nout) * sizeof(u64); }
extern "C" int evab_lazy_rotsum(evab_ctx *c, int ell, int nout, uint64_t *o, const uint64_t *a, const uint64_t *ext, int n, const uint64_t *elts,
                                const uint64_t *const *keys, const uint64_t *const *cadds, const uint64_t *const *wts, void *work, void *stream) {
  if (n < 1 || n > LRS_MAX) return fail("evab_lazy_rotsum: 1..16 rotations");
  const u32 *perms[LRS_MAX];
  for (int i = 0; i < n; i++) {
    u32 *perm = nullptr, *ctab = nullptr;
    if (galois_tables(c, elts[i], &perm, &ctab, "evab_lazy_rotsum")) return 1;
    perms[i] = perm;
  }
  BE_BEGIN return lazy_rotsum_impl(be, c->v, ell, nout, o, a, ext, n, perms, keys, cadds, wts, (u64 *)work);
}
extern "C" int evab_memset_zero(evab_ctx *c, void *d, size_t bytes, void *stream) {
  CUDA_OK(cudaSetDevice(c->device));
  CUDA_OK(cudaMemsetAsync(d, 0, bytes, S(stream)));
  return 0;
}
extern "C" int evab_rotate(evab_ctx *c, int ell, uint64_t *o, const uint64_t *a, uint64_t elt, const uint64_t *key, void *work, void *stream) {
  u32 *perm = nullptr;
  {
    std::lock_guard<std::mutex> g(c->mu);
    auto it = c->perms.find(elt);
    if (it == c->perms.end()) return fail("evab_rotate: call evab_galois_prepare(elt) first");
    perm = it->second;
  }
  BE_BEGIN return rotate_impl(be, c->v, ell, o, a, perm, key, (u64 *)work);
}
