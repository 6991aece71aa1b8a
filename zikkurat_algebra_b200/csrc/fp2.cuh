// Quadratic extension Fp2 = Fp[u]/(u^2 + 1) for the G2 groups of BN254 and BLS12-381 (scope row 8f.3).
//
// Same representation as the reference (c0 || c1, both Montgomery Fp; u^2 = -1:
// lib/cbits/curves/fields/mont/bn128_Fp2_mont.c:182-196).  An element is an Fe<Ext2<B>> with 2*B::L limbs, so
// every generic piece of the MSM (memory images, shuffles, selects, zero tests, the group law in ec.cuh) works
// unchanged; only the arithmetic below is specialised.
//   mul : c0 = a0*b0 + a1*(-b1),  c1 = a0*b1 + a1*b0     two fused two-product reductions (fe_mul2)
//   sqr : c0 = (a0+a1)*(a0-a1),   c1 = (2*a0)*a1          two base multiplications
// On the device the base products are always the out-of-line versions (fe_mul_call / fe_mul2_call), which
// keeps the G2 kernels' code size and compile time in check.
#pragma once
#include "fp.cuh"

namespace zk {

template <class B>
struct Ext2 {
  using Base = B;
  static constexpr int L = 2 * B::L;
  static constexpr bool THREE_MOD_FITS = B::THREE_MOD_FITS;
  ZK_HD static constexpr uint32_t one(int i) { return i < B::L ? B::one(i) : 0u; }
};

template <class B>
ZK_HD Fe<B> ext_lo(const Fe<Ext2<B>>& a) {
  Fe<B> r;
#pragma unroll
  for (int i = 0; i < B::L; i++) r.l[i] = a.l[i];
  return r;
}
template <class B>
ZK_HD Fe<B> ext_hi(const Fe<Ext2<B>>& a) {
  Fe<B> r;
#pragma unroll
  for (int i = 0; i < B::L; i++) r.l[i] = a.l[B::L + i];
  return r;
}
template <class B>
ZK_HD Fe<Ext2<B>> ext_make(const Fe<B>& lo, const Fe<B>& hi) {
  Fe<Ext2<B>> r;
#pragma unroll
  for (int i = 0; i < B::L; i++) { r.l[i] = lo.l[i]; r.l[B::L + i] = hi.l[i]; }
  return r;
}

template <class B>
ZK_HD Fe<Ext2<B>> ext2_mul(const Fe<Ext2<B>>& a, const Fe<Ext2<B>>& b) {
  Fe<B> a0 = ext_lo<B>(a), a1 = ext_hi<B>(a), b0 = ext_lo<B>(b), b1 = ext_hi<B>(b);
  return ext_make<B>(fe_mul2_call<B>(a0, b0, a1, fe_neg<B>(b1)), fe_mul2_call<B>(a0, b1, a1, b0));
}
template <class B>
ZK_HD Fe<Ext2<B>> ext2_sqr(const Fe<Ext2<B>>& a) {
  Fe<B> a0 = ext_lo<B>(a), a1 = ext_hi<B>(a);
  return ext_make<B>(fe_mul_call<B>(fe_add<B>(a0, a1), fe_sub<B>(a0, a1)), fe_mul_call<B>(fe_dbl<B>(a0), a1));
}
template <class B>
ZK_HD Fe<Ext2<B>> ext2_inv(const Fe<Ext2<B>>& a) {  // conj(a) / (a0^2 + a1^2)
  Fe<B> a0 = ext_lo<B>(a), a1 = ext_hi<B>(a);
  Fe<B> ni = fe_inv<B>(fe_mul2_call<B>(a0, a0, a1, a1));
  return ext_make<B>(fe_mul_call<B>(a0, ni), fe_neg<B>(fe_mul_call<B>(a1, ni)));
}

// Full specialisations of the generic field interface for one concrete base field.
#define ZK_DEFINE_EXT2(B)                                                                                              \
  template <> ZK_HD Fe<Ext2<B>> fe_mul<Ext2<B>>(const Fe<Ext2<B>>& a, const Fe<Ext2<B>>& b) { return ext2_mul<B>(a, b); } \
  template <> ZK_HD Fe<Ext2<B>> fe_sqr<Ext2<B>>(const Fe<Ext2<B>>& a) { return ext2_sqr<B>(a); }                        \
  template <> ZK_HD Fe<Ext2<B>> fe_add<Ext2<B>>(const Fe<Ext2<B>>& a, const Fe<Ext2<B>>& b) {                            \
    return ext_make<B>(fe_add<B>(ext_lo<B>(a), ext_lo<B>(b)), fe_add<B>(ext_hi<B>(a), ext_hi<B>(b))); }                  \
  template <> ZK_HD Fe<Ext2<B>> fe_sub<Ext2<B>>(const Fe<Ext2<B>>& a, const Fe<Ext2<B>>& b) {                            \
    return ext_make<B>(fe_sub<B>(ext_lo<B>(a), ext_lo<B>(b)), fe_sub<B>(ext_hi<B>(a), ext_hi<B>(b))); }                  \
  template <> ZK_HD Fe<Ext2<B>> fe_neg<Ext2<B>>(const Fe<Ext2<B>>& a) {                                                  \
    return ext_make<B>(fe_neg<B>(ext_lo<B>(a)), fe_neg<B>(ext_hi<B>(a))); }                                              \
  template <> ZK_HD Fe<Ext2<B>> fe_mul2<Ext2<B>>(const Fe<Ext2<B>>& a, const Fe<Ext2<B>>& b, const Fe<Ext2<B>>& c,       \
                                                 const Fe<Ext2<B>>& d) {                                                 \
    return fe_add<Ext2<B>>(ext2_mul<B>(a, b), ext2_mul<B>(c, d)); }                                                      \
  template <> ZK_HD Fe<Ext2<B>> fe_inv<Ext2<B>>(const Fe<Ext2<B>>& a) { return ext2_inv<B>(a); }                        \
  template <> ZK_HD Fe<Ext2<B>> fe_mul_call<Ext2<B>>(Fe<Ext2<B>> a, Fe<Ext2<B>> b) { return ext2_mul<B>(a, b); }        \
  template <> ZK_HD Fe<Ext2<B>> fe_sqr_call<Ext2<B>>(Fe<Ext2<B>> a) { return ext2_sqr<B>(a); }                          \
  template <> ZK_HD Fe<Ext2<B>> fe_mul2_call<Ext2<B>>(Fe<Ext2<B>> a, Fe<Ext2<B>> b, Fe<Ext2<B>> c, Fe<Ext2<B>> d) {      \
    return fe_add<Ext2<B>>(ext2_mul<B>(a, b), ext2_mul<B>(c, d)); }

}  // namespace zk
