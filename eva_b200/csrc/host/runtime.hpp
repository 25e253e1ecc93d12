// runtime.hpp -- C++ client of the C-ABI (include/evab200.h): RAII device
// context, buffers, host copies of the moduli.  Everything the host layer does
// on the GPU goes through the extern "C" boundary, exactly as a reference-side
// binding would.
#pragma once
#include "../../../include/evab200.h"
#include <cstdint>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

namespace evab {

typedef std::uint64_t u64;

inline void check(int rc) {
  if (rc) throw std::runtime_error(std::string("evab200: ") + evab_last_error());
}

// host-side modular helpers (parameter setup and client-side code only)
namespace hmod {
typedef unsigned __int128 u128;
inline u64 mul(u64 a, u64 b, u64 p) { return (u64)(((u128)a * b) % p); }
inline u64 pow(u64 a, u64 e, u64 p) { u64 r = 1 % p; a %= p; while (e) { if (e & 1) r = mul(r, a, p); a = mul(a, a, p); e >>= 1; } return r; }
inline u64 inv(u64 a, u64 p) { return pow(a, p - 2, p); }
inline bool isPrime(u64 n) {
  static const u64 bases[12] = {2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37};
  if (n < 2) return false;
  for (u64 b : bases) { if (n == b) return true; if (n % b == 0) return false; }
  u64 d = n - 1; int r = 0;
  while (!(d & 1)) { d >>= 1; r++; }
  for (u64 b : bases) {
    u64 x = pow(b, d, n);
    if (x == 1 || x == n - 1) continue;
    bool comp = true;
    for (int j = 1; j < r && comp; j++) { x = mul(x, x, n); if (x == n - 1) comp = false; }
    if (comp) return false;
  }
  return true;
}
// seal::CoeffModulus::Create(N, bit_sizes) order (reference eva/seal/seal.cpp:181-182;
// SURVEY Appendix A.1): per bit size scan 2^b - 2N + 1 downwards in steps of 2N,
// then each entry takes the smallest unused prime of its size.
inline std::vector<u64> createCoeffModulus(u64 N, const std::vector<int> &bits) {
  std::map<int, std::vector<u64>> table;
  std::map<int, int> count;
  for (int b : bits) { if (b < 2 || b > 60) throw std::invalid_argument("bit_sizes is invalid"); count[b]++; }
  const u64 factor = 2 * N;
  for (auto &e : count) {
    const int b = e.first;
    u64 value = ((u64(1) << b) - 1) / factor * factor + 1, lower = u64(1) << (b - 1);
    auto &v = table[b];
    while ((int)v.size() < e.second && value > lower) { if (isPrime(value)) v.push_back(value); value -= factor; }
    if ((int)v.size() < e.second) throw std::logic_error("failed to find enough qualifying primes");
  }
  std::vector<u64> out;
  for (int b : bits) { out.push_back(table[b].back()); table[b].pop_back(); }
  return out;
}
}  // namespace hmod

class Device {
public:
  Device(u64 N, const std::vector<u64> &primes, int device) : N_(N), primes_(primes), device_(device) {
    check(evab_ctx_create(N, primes.data(), (int)primes.size(), device, &ctx_));
  }
  ~Device() { evab_ctx_destroy(ctx_); }
  Device(const Device &) = delete;
  evab_ctx *ctx() const { return ctx_; }
  u64 N() const { return N_; }
  int k() const { return (int)primes_.size(); }
  int device() const { return device_; }
  const std::vector<u64> &primes() const { return primes_; }

  void *alloc(std::size_t bytes) { void *p = nullptr; check(evab_malloc(ctx_, bytes, &p, nullptr)); return p; }
  void free(void *p) { if (p) evab_free(ctx_, p, nullptr); }
  void upload(void *d, const void *h, std::size_t bytes, void *stream = nullptr) { check(evab_upload(ctx_, d, h, bytes, stream)); }
  void download(void *h, const void *d, std::size_t bytes, void *stream = nullptr) { check(evab_download(ctx_, h, d, bytes, stream)); }
  void sync(void *stream = nullptr) { check(evab_sync(ctx_, stream)); }

private:
  u64 N_;
  std::vector<u64> primes_;
  int device_;
  evab_ctx *ctx_ = nullptr;
};

// owning device buffer of u64 words
// Page-locked host memory for the buffers that cross the boundary (HostCipher / HostPlain):
// size-class free lists so that steady-state execute() calls never hit cudaHostAlloc.
class PinnedPool {
public:
  static PinnedPool &get() { static PinnedPool p; return p; }
  void *alloc(std::size_t bytes) {
    const std::size_t cls = sizeClass(bytes);
    {
      std::lock_guard<std::mutex> g(mu_);
      auto &fl = free_[cls];
      if (!fl.empty()) { void *p = fl.back(); fl.pop_back(); return p; }
    }
    void *p = nullptr;
    check(evab_host_alloc((std::size_t)1 << cls, &p));
    return p;
  }
  void release(void *p, std::size_t bytes) {
    std::lock_guard<std::mutex> g(mu_);
    free_[sizeClass(bytes)].push_back(p);
  }
private:
  static std::size_t sizeClass(std::size_t bytes) { std::size_t c = 12; while (((std::size_t)1 << c) < bytes) c++; return c; }
  std::mutex mu_;
  std::map<std::size_t, std::vector<void *>> free_;   // blocks stay pinned for the life of the process
};
template <class T> struct PinnedAlloc {
  typedef T value_type;
  PinnedAlloc() {}
  template <class U> PinnedAlloc(const PinnedAlloc<U> &) {}
  T *allocate(std::size_t n) { return static_cast<T *>(PinnedPool::get().alloc(n * sizeof(T))); }
  void deallocate(T *p, std::size_t n) { PinnedPool::get().release(p, n * sizeof(T)); }
  // resize() of a buffer that is about to be overwritten by a download: no zero fill
  template <class U> void construct(U *) {}
  template <class U, class... A> void construct(U *p, A &&...a) { ::new ((void *)p) U(std::forward<A>(a)...); }
  template <class U> bool operator==(const PinnedAlloc<U> &) const { return true; }
  template <class U> bool operator!=(const PinnedAlloc<U> &) const { return false; }
};
typedef std::vector<u64, PinnedAlloc<u64>> HostBuf;

class DBuf {
public:
  DBuf() {}
  DBuf(std::shared_ptr<Device> dev, std::size_t words) : dev_(std::move(dev)), words_(words) { p_ = (u64 *)dev_->alloc(words * 8); }
  ~DBuf() { if (dev_) dev_->free(p_); }
  DBuf(DBuf &&o) noexcept { *this = std::move(o); }
  DBuf &operator=(DBuf &&o) noexcept { std::swap(dev_, o.dev_); std::swap(p_, o.p_); std::swap(words_, o.words_); return *this; }
  DBuf(const DBuf &) = delete;
  u64 *get() const { return p_; }
  std::size_t words() const { return words_; }
  explicit operator bool() const { return p_ != nullptr; }
private:
  std::shared_ptr<Device> dev_;
  u64 *p_ = nullptr;
  std::size_t words_ = 0;
};

}  // namespace evab
