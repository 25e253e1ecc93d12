// csprng.hpp -- the randomness of key generation and encryption (secret key, encryption masks, error
// polynomials, the uniform halves of the public / relinearization / Galois keys).
//
// The reference samples from SEAL's Blake2/SHAKE CSPRNG seeded with 512 bits.  This backend uses the ChaCha20
// block function (RFC 8439) as a stream generator keyed with 256 bits from the operating system (getrandom(2),
// /dev/urandom as a fallback).  A non-zero `seed` selects a DETERMINISTIC stream derived from that 64-bit value:
// it exists for reproducible tests and fixtures only and carries at most 64 bits of security.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#if defined(__linux__)
#include <sys/random.h>
#endif

namespace evab {

class ChaChaRng {
public:
  explicit ChaChaRng(std::uint64_t seed) {
    std::uint32_t key[8];
    if (seed == 0) {
      osEntropy(key, sizeof(key));
      deterministic_ = false;
    } else {   // test-only: SplitMix64 expansion of the seed
      std::uint64_t x = seed;
      for (int i = 0; i < 4; i++) {
        x += 0x9E3779B97F4A7C15ull;
        std::uint64_t z = x;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; z ^= z >> 31;
        key[2 * i] = (std::uint32_t)z; key[2 * i + 1] = (std::uint32_t)(z >> 32);
      }
      deterministic_ = true;
    }
    static const std::uint32_t sigma[4] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u};   // "expand 32-byte k"
    std::memcpy(state_, sigma, 16);
    std::memcpy(state_ + 4, key, 32);
    state_[12] = state_[13] = state_[14] = state_[15] = 0;   // 64-bit block counter, 64-bit nonce 0 (one stream per key)
    pos_ = 16;
    std::memset(key, 0, sizeof(key));
  }
  bool deterministic() const { return deterministic_; }
  std::uint64_t operator()() {
    if (pos_ + 2 > 16) refill();
    const std::uint64_t v = (std::uint64_t)block_[pos_] | ((std::uint64_t)block_[pos_ + 1] << 32);
    pos_ += 2;
    return v;
  }
  // uniform in [0, bound) by rejection (no modulo bias)
  std::uint64_t below(std::uint64_t bound) {
    const std::uint64_t lim = ~0ull - (~0ull % bound + 1) % bound;   // largest multiple of bound, minus one
    std::uint64_t v;
    do { v = (*this)(); } while (v > lim);
    return v % bound;
  }

private:
  static std::uint32_t rotl(std::uint32_t x, int n) { return (x << n) | (x >> (32 - n)); }
  static void qr(std::uint32_t *s, int a, int b, int c, int d) {
    s[a] += s[b]; s[d] ^= s[a]; s[d] = rotl(s[d], 16);
    s[c] += s[d]; s[b] ^= s[c]; s[b] = rotl(s[b], 12);
    s[a] += s[b]; s[d] ^= s[a]; s[d] = rotl(s[d], 8);
    s[c] += s[d]; s[b] ^= s[c]; s[b] = rotl(s[b], 7);
  }
  void refill() {
    std::uint32_t x[16];
    std::memcpy(x, state_, 64);
    for (int i = 0; i < 10; i++) {
      qr(x, 0, 4, 8, 12); qr(x, 1, 5, 9, 13); qr(x, 2, 6, 10, 14); qr(x, 3, 7, 11, 15);
      qr(x, 0, 5, 10, 15); qr(x, 1, 6, 11, 12); qr(x, 2, 7, 8, 13); qr(x, 3, 4, 9, 14);
    }
    for (int i = 0; i < 16; i++) block_[i] = x[i] + state_[i];
    if (++state_[12] == 0) ++state_[13];
    pos_ = 0;
  }
  static void osEntropy(void *buf, std::size_t n) {
#if defined(__linux__)
    std::size_t got = 0;
    while (got < n) {
      const ssize_t r = getrandom((char *)buf + got, n - got, 0);
      if (r <= 0) break;
      got += (std::size_t)r;
    }
    if (got == n) return;
#endif
    std::FILE *f = std::fopen("/dev/urandom", "rb");
    if (!f || std::fread(buf, 1, n, f) != n) { if (f) std::fclose(f); throw std::runtime_error("no entropy source (getrandom / /dev/urandom)"); }
    std::fclose(f);
  }
  std::uint32_t state_[16], block_[16];
  int pos_;
  bool deterministic_;
};

}  // namespace evab
