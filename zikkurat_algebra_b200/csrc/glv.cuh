// GLV scalar decomposition for the j = 0 curves BN254 and BLS12-381 (G1).
//
// phi(x, y) = (beta x, y) is an endomorphism of y^2 = x^3 + b that acts as multiplication by lambda on the prime-order
// subgroup, lambda^2 + lambda + 1 = 0 (mod r).  Every scalar is split as  k = k1 + k2 lambda (mod r)  with
// |k1|, |k2| < 2^127, so the MSM  sum k_i P_i  becomes an MSM over the 2n points  P_i, phi(P_i)  with 127-bit
// scalars: HALF the windows -- half the buckets to reduce and half of the 255 serial doublings of the window
// combination, which is what a caller of a single MSM ends up waiting for -- at the same number of bucket insertions.
// The group element is the same, hence the canonical affine output stays bit-identical to the reference's
//   lib/cbits/curves/g1/proj/bn128_G1_proj.c:506-586  (which walks all 254/255 bits).
// The reference lists (beta, lambda) for both curves (codegen/src/Zikkurat/CodeGen/Curve/Params.hs:162-165,200-203) but
// does not use them in its MSM.  Constants: tools/gen_params.py (glv()) -> curve_params.cuh.
//
// The points only have to lie in the subgroup for the split to be an identity, which is the MSM's own domain (G1
// elements of the SRS; the reference does not validate its inputs either).  $ZKB200_GLV=0 switches the split off.
#pragma once
#include "curve_params.cuh"
#include "fp.cuh"

namespace zk {

template <class C> struct GlvOf { static constexpr bool available = false; };
template <> struct GlvOf<Bn254> { static constexpr bool available = true; using type = Bn254Glv; };
template <> struct GlvOf<Bls12381> { static constexpr bool available = true; using type = Bls12381Glv; };

// c = (k * g + 2^(SHIFT-1)) >> SHIFT : k 8 limbs, g 7 limbs, SHIFT = 320 -> 5 limbs (limbs 10..14 of the product)
template <class G, int WHICH>
ZK_HD void glv_quotient(const uint32_t* k, uint32_t* c) {
  uint32_t prod[15];
#pragma unroll
  for (int i = 0; i < 15; i++) prod[i] = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    uint64_t carry = 0;
#pragma unroll
    for (int j = 0; j < 7; j++) {
      const uint32_t g = WHICH == 1 ? G::g1(j) : G::g2(j);
      uint64_t t = (uint64_t)k[i] * g + prod[i + j] + carry;   // <= (2^32-1)^2 + 2(2^32-1) = 2^64 - 1
      prod[i + j] = (uint32_t)t;
      carry = t >> 32;
    }
    prod[i + 7] = (uint32_t)carry;
  }
  // + 2^319 = bit 31 of limb 9, carried upwards
  uint64_t t = (uint64_t)prod[9] + 0x80000000u;
  uint32_t carry = (uint32_t)(t >> 32);
#pragma unroll
  for (int i = 10; i < 15; i++) {
    uint64_t u = (uint64_t)prod[i] + carry;
    c[i - 10] = (uint32_t)u;
    carry = (uint32_t)(u >> 32);
  }
}

// acc (5 limbs, two's complement mod 2^160) += sign * c (5 limbs) * m (4 limbs)
template <int SIGN, class M>
ZK_HD void glv_addmul(uint32_t* acc, const uint32_t* c, M m) {
  uint32_t t[5] = {0, 0, 0, 0, 0};
#pragma unroll
  for (int i = 0; i < 5; i++) {
    uint64_t carry = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      if (i + j < 5) {
        uint64_t v = (uint64_t)c[i] * m(j) + t[i + j] + carry;
        t[i + j] = (uint32_t)v;
        carry = v >> 32;
      }
    }
    if (i == 0) t[4] = (uint32_t)carry;   // rows i >= 1 carry out beyond 2^160: dropped (arithmetic mod 2^160)
  }
  if (SIGN > 0) {
    uint64_t carry = 0;
#pragma unroll
    for (int i = 0; i < 5; i++) { uint64_t v = (uint64_t)acc[i] + t[i] + carry; acc[i] = (uint32_t)v; carry = v >> 32; }
  } else {
    uint64_t borrow = 0;
#pragma unroll
    for (int i = 0; i < 5; i++) { uint64_t v = (uint64_t)acc[i] - t[i] - borrow; acc[i] = (uint32_t)v; borrow = (v >> 32) & 1u; }
  }
}

// sign + magnitude (4 limbs) of a 160-bit two's complement value known to be < 2^127 in absolute value
ZK_HD bool glv_abs(uint32_t* acc, uint32_t* mag) {
  const bool neg = (acc[4] >> 31) != 0;
  if (neg) {
    uint64_t carry = 1;
#pragma unroll
    for (int i = 0; i < 5; i++) { uint64_t v = (uint64_t)(~acc[i]) + carry; acc[i] = (uint32_t)v; carry = v >> 32; }
  }
#pragma unroll
  for (int i = 0; i < 4; i++) mag[i] = acc[i];
  return neg;
}

template <class G> struct GlvRow {
  struct A1 { ZK_HD uint32_t operator()(int j) const { return G::a1(j); } };
  struct B1 { ZK_HD uint32_t operator()(int j) const { return G::b1(j); } };
  struct A2 { ZK_HD uint32_t operator()(int j) const { return G::a2(j); } };
  struct B2 { ZK_HD uint32_t operator()(int j) const { return G::b2(j); } };
};

// k (8 limbs, ANY value < 2^256: std_coeff scalars are not reduced) -> |k1|, |k2| (4 limbs each, < 2^127) and their signs
template <class G>
ZK_HD void glv_decompose(const uint32_t* k, uint32_t* k1, bool& neg1, uint32_t* k2, bool& neg2) {
  uint32_t c1[5], c2[5];
  glv_quotient<G, 1>(k, c1);
  glv_quotient<G, 2>(k, c2);
  uint32_t a[5] = {k[0], k[1], k[2], k[3], k[4]};
  glv_addmul<G::S11>(a, c1, typename GlvRow<G>::A1());
  glv_addmul<G::S12>(a, c2, typename GlvRow<G>::A2());
  neg1 = glv_abs(a, k1);
  uint32_t b[5] = {0, 0, 0, 0, 0};
  glv_addmul<G::S21>(b, c1, typename GlvRow<G>::B1());
  glv_addmul<G::S22>(b, c2, typename GlvRow<G>::B2());
  neg2 = glv_abs(b, k2);
}

// phi(P) for an affine Montgomery point: x -> beta x (infinity, the all-0xFF record, is kept as it is by the caller)
template <class P, class G>
ZK_HD Fe<P> glv_beta_x(const Fe<P>& x) {
  Fe<P> b;
#pragma unroll
  for (int i = 0; i < P::L; i++) b.l[i] = G::beta(i);
  return fe_mul<P>(x, b);
}

}  // namespace zk
