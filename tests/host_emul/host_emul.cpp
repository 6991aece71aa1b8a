// TEST INFRASTRUCTURE ONLY.  Compiles the device headers (fp.cuh / ec.cuh / recode.cuh) with the HOST
// compiler, where hd.cuh emulates the PTX carry-chain primitives, so the limb-level algorithms can be
// checked against the oracles on a machine without a GPU.  Never linked into the product library.
#include <stdint.h>
#include <string.h>

#include "curve_params.cuh"
#include "ec.cuh"
#include "recode.cuh"

using namespace zk;

template <class P>
static Fe<P> ld(const uint64_t* s) {
  Fe<P> r;
  memcpy(r.l, s, 4 * P::L);
  return r;
}
template <class P>
static void st(uint64_t* d, const Fe<P>& a) {
  memcpy(d, a.l, 4 * P::L);
}
template <class P>
static Affine<P> ldaff(const uint64_t* s) {
  Affine<P> r;
  r.x = ld<P>(s);
  r.y = ld<P>(s + P::L / 2);
  return r;
}
template <class P>
static void staff(uint64_t* d, const Xyzz<P>& a) {
  Affine<P> o;
  if (!xyzz_to_affine<P>(a, o)) { memset(d, 0xff, 8 * P::L); return; }
  st<P>(d, o.x);
  st<P>(d + P::L / 2, o.y);
}
template <class P>
static Xyzz<P> sum_list(long n, const uint64_t* pts, const uint8_t* neg) {
  Xyzz<P> acc = xyzz_inf<P>();
  for (long i = 0; i < n; i++) {
    Affine<P> p = ldaff<P>(pts + (size_t)i * P::L);
    if (affine_is_inf<P>(p)) continue;
    if (neg && neg[i]) p = affine_neg<P>(p);
    xyzz_madd<P>(acc, p);
  }
  return acc;
}

#define DEFINE(NAME, C)                                                                                         \
  extern "C" void he_##NAME##_fp_mul(const uint64_t* a, const uint64_t* b, uint64_t* t) {                       \
    st<C::Fp>(t, fe_mul<C::Fp>(ld<C::Fp>(a), ld<C::Fp>(b))); }                                                  \
  extern "C" void he_##NAME##_fp_mul2(const uint64_t* a, const uint64_t* b, const uint64_t* c, const uint64_t* d, uint64_t* t) { \
    st<C::Fp>(t, fe_mul2<C::Fp>(ld<C::Fp>(a), ld<C::Fp>(b), ld<C::Fp>(c), ld<C::Fp>(d))); }                       \
  extern "C" void he_##NAME##_fp_sqr(const uint64_t* a, uint64_t* t) { st<C::Fp>(t, fe_sqr<C::Fp>(ld<C::Fp>(a))); }  \
  extern "C" void he_##NAME##_fp_add(const uint64_t* a, const uint64_t* b, uint64_t* t) {                       \
    st<C::Fp>(t, fe_add<C::Fp>(ld<C::Fp>(a), ld<C::Fp>(b))); }                                                  \
  extern "C" void he_##NAME##_fp_sub(const uint64_t* a, const uint64_t* b, uint64_t* t) {                       \
    st<C::Fp>(t, fe_sub<C::Fp>(ld<C::Fp>(a), ld<C::Fp>(b))); }                                                  \
  extern "C" void he_##NAME##_fp_neg(const uint64_t* a, uint64_t* t) { st<C::Fp>(t, fe_neg<C::Fp>(ld<C::Fp>(a))); } \
  extern "C" void he_##NAME##_fp_inv(const uint64_t* a, uint64_t* t) { st<C::Fp>(t, fe_inv<C::Fp>(ld<C::Fp>(a))); } \
  extern "C" void he_##NAME##_fp_inv_fermat(const uint64_t* a, uint64_t* t) { st<C::Fp>(t, fe_inv_fermat<C::Fp>(ld<C::Fp>(a))); } \
  extern "C" void he_##NAME##_fr_sqr(const uint64_t* a, uint64_t* t) { st<C::Fr>(t, fe_sqr<C::Fr>(ld<C::Fr>(a))); }       \
  extern "C" void he_##NAME##_fr_mul(const uint64_t* a, const uint64_t* b, uint64_t* t) {                       \
    st<C::Fr>(t, fe_mul<C::Fr>(ld<C::Fr>(a), ld<C::Fr>(b))); }                                                  \
  extern "C" void he_##NAME##_fr_to_std(const uint64_t* a, uint64_t* t) {                                       \
    st<C::Fr>(t, fe_from_mont<C::Fr>(ld<C::Fr>(a))); }                                                          \
  extern "C" void he_##NAME##_sum_list(long n, const uint64_t* pts, const uint8_t* neg, uint64_t* t_aff) {      \
    staff<C::Fp>(t_aff, sum_list<C::Fp>(n, pts, neg)); }                                                        \
  extern "C" void he_##NAME##_add_lists(long n1, const uint64_t* p1, long n2, const uint64_t* p2,               \
                                        uint64_t* t_aff, uint64_t* t_proj, uint64_t* t_jac) {                   \
    Xyzz<C::Fp> s = xyzz_add<C::Fp>(sum_list<C::Fp>(n1, p1, 0), sum_list<C::Fp>(n2, p2, 0));                     \
    staff<C::Fp>(t_aff, s);                                                                                     \
    Fe<C::Fp> X, Y, Z;                                                                                          \
    xyzz_to_proj<C::Fp>(s, X, Y, Z);                                                                            \
    st<C::Fp>(t_proj, X); st<C::Fp>(t_proj + C::Fp::L / 2, Y); st<C::Fp>(t_proj + C::Fp::L, Z);                 \
    Xyzz<C::Fp> back = xyzz_from_proj<C::Fp>(X, Y, Z);                                                          \
    xyzz_to_jac<C::Fp>(back, X, Y, Z);                                                                          \
    st<C::Fp>(t_jac, X); st<C::Fp>(t_jac + C::Fp::L / 2, Y); st<C::Fp>(t_jac + C::Fp::L, Z);                    \
    Xyzz<C::Fp> back2 = xyzz_from_jac<C::Fp>(X, Y, Z);                                                          \
    Xyzz<C::Fp> d = xyzz_dbl<C::Fp>(back2);                                                                     \
    (void)d;                                                                                                    \
  }                                                                                                             \
  extern "C" void he_##NAME##_dbl_list(long n, const uint64_t* pts, uint64_t* t_aff) {                          \
    staff<C::Fp>(t_aff, xyzz_dbl<C::Fp>(sum_list<C::Fp>(n, pts, 0))); }

DEFINE(bn128, Bn254)
DEFINE(bls12_381, Bls12381)

// signed-digit recoding of one scalar: out_key[w], out_neg[w] for w < nwin
extern "C" void he_recode(const uint64_t* scalar, int nbits, int c, int nwin, uint32_t* keys, uint32_t* negs) {
  uint32_t limbs[8];
  memcpy(limbs, scalar, 32);
  uint32_t carry = 0;
  for (int w = 0; w < nwin; w++) {
    uint32_t key, neg;
    recode_digit(limbs, nbits, c, w, carry, key, neg);
    keys[w] = key;
    negs[w] = neg;
  }
  keys[nwin] = carry;  // must be 0 when nwin*c >= nbits+1
}
