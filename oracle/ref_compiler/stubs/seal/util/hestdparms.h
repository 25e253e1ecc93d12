// HomomorphicEncryption.org standard bit budgets, as in SEAL 3.6
// seal/util/hestdparms.h [SEAL-KNOWLEDGE]; called at
// /root/reference/eva/ckks/ckks_compiler.h:176-193.
#pragma once
#include <cstddef>
namespace seal { namespace util {
#define EVA_HESTD(NAME, a, b, c, d, e, f)                                   \
  inline int NAME(std::size_t n) {                                       \
    return n == 1024 ? a : n == 2048 ? b : n == 4096 ? c : n == 8192 ? d    \
         : n == 16384 ? e : n == 32768 ? f : 0;                             \
  }
EVA_HESTD(seal_he_std_parms_128_tc, 27, 54, 109, 218, 438, 881)
EVA_HESTD(seal_he_std_parms_192_tc, 19, 37, 75, 152, 305, 611)
EVA_HESTD(seal_he_std_parms_256_tc, 14, 29, 58, 118, 237, 476)
EVA_HESTD(seal_he_std_parms_128_tq, 25, 51, 101, 202, 411, 827)
EVA_HESTD(seal_he_std_parms_192_tq, 17, 35, 70, 141, 284, 571)
EVA_HESTD(seal_he_std_parms_256_tq, 13, 27, 54, 109, 220, 443)
} }
