// TEST INFRASTRUCTURE ONLY (oracle).  Corrected replacement for the reference's
// lib/cbits/platform.h, placed FIRST on the -I path when the unmodified reference
// sources are compiled into oracle/_ref/libzk_ref.so.
//
// Why: the shipped header is broken in both #ifdef branches
//   /root/reference/lib/cbits/platform.h:13,17   (ARCH_X86_64: passes *tgt where a pointer is needed)
//   /root/reference/lib/cbits/platform.h:51-52   (portable: stores an uninitialised word)
// Semantics fixed by the reference's own test-suite:
//   /root/reference/test/src/ZK/Test/Platform/Properties.hs:74-97
#pragma once
#include <stdint.h>

static inline uint8_t addcarry_u64(uint8_t c, uint64_t a, uint64_t b, uint64_t *t) {
  unsigned __int128 s = (unsigned __int128)a + b + c;
  *t = (uint64_t)s;
  return (uint8_t)(s >> 64);
}

static inline uint8_t subborrow_u64(uint8_t c, uint64_t a, uint64_t b, uint64_t *t) {
  unsigned __int128 s = (unsigned __int128)a - b - c;
  *t = (uint64_t)s;
  return (uint8_t)((s >> 64) & 1);
}

static inline uint8_t addcarry_u128_inplace(uint64_t *lo, uint64_t *hi, uint64_t alo, uint64_t ahi) {
  uint8_t c = addcarry_u64(0, *lo, alo, lo);
  return addcarry_u64(c, *hi, ahi, hi);
}
