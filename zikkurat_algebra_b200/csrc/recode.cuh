// Signed-digit window recoding of a 256-bit scalar (8 x 32-bit limbs, little endian).
//
// Replaces the reference's per-window digit extraction
//   lib/cbits/curves/g1/proj/bn128_G1_proj.c:520-538,551-556   (unsigned c-bit digits)
// with signed digits d_w in [-2^(c-1), 2^(c-1)] so that only 2^(c-1) buckets per window are needed:
//   k = sum_w d_w * 2^(c*w),   bucket key = |d_w| (0 = no insertion), sign -> negate the point.
// The recoding is exact for every 256-bit integer (the std_coeff entry points accept un-reduced
// scalars) as long as  nwin * c >= nbits + 1.
#pragma once
#include "hd.cuh"

namespace zk {

// bits [lo, lo+c) of the scalar, zero beyond nbits; c <= 31
ZK_HD uint32_t scalar_bits(const uint32_t* limbs, int nbits, int lo, int c) {
  if (lo >= nbits) return 0;
  int w = lo >> 5, s = lo & 31;
  uint64_t v = limbs[w];
  if (w + 1 < 8) v |= (uint64_t)limbs[w + 1] << 32;
  uint32_t d = (uint32_t)(v >> s) & ((1u << c) - 1u);
  int over = lo + c - nbits;  // bits past the declared length are ignored
  if (over > 0) d &= (1u << (c - over)) - 1u;
  return d;
}

// window w (must be visited in increasing w, threading `carry` through)
ZK_HD void recode_digit(const uint32_t* limbs, int nbits, int c, int w, uint32_t& carry, uint32_t& key, uint32_t& neg) {
  uint32_t raw = scalar_bits(limbs, nbits, w * c, c) + carry;
  uint32_t half = 1u << (c - 1);
  if (raw > half) {
    key = (1u << c) - raw;  // |raw - 2^c|, in [0, 2^(c-1))
    neg = key != 0;
    carry = 1;
  } else {
    key = raw;
    neg = 0;
    carry = 0;
  }
}

// number of signed windows needed for scalars of `nbits` bits
ZK_HD constexpr int signed_windows(int nbits, int c) { return (nbits + 1 + c - 1) / c; }

}  // namespace zk
