// TEST INFRASTRUCTURE ONLY.  Compiles the device headers (fp.cuh / ec.cuh / recode.cuh) with the HOST
// compiler, where hd.cuh emulates the PTX carry-chain primitives, so the limb-level algorithms can be
// checked against the oracles on a machine without a GPU.  Never linked into the product library.
#include <stdint.h>
#include <string.h>

#include "curve_params.cuh"
#include "ec.cuh"
#include "recode.cuh"
#include "aff_plan.cuh"
#include "glv.cuh"
#include "red_plan.cuh"
#include <vector>

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
  extern "C" void he_##NAME##_fp_mul_pair(const uint64_t* a, const uint64_t* b, const uint64_t* c, uint64_t* t1, uint64_t* t2) { \
    FePair<C::Fp> r = fe_mul_pair_call<C::Fp>(ld<C::Fp>(a), ld<C::Fp>(b), ld<C::Fp>(c)); st<C::Fp>(t1, r.u); st<C::Fp>(t2, r.v); } \
  extern "C" void he_##NAME##_fp_mul_kara(const uint64_t* a, const uint64_t* b, uint64_t* t) {                  \
    Fe<C::Fp> x = ld<C::Fp>(a), y = ld<C::Fp>(b), r; mont_mul_kara_limbs<C::Fp>(r.l, x.l, y.l); st<C::Fp>(t, r); }  \
  extern "C" void he_##NAME##_fp_sqr(const uint64_t* a, uint64_t* t) { st<C::Fp>(t, fe_sqr<C::Fp>(ld<C::Fp>(a))); }  \
  extern "C" void he_##NAME##_fp_add(const uint64_t* a, const uint64_t* b, uint64_t* t) {                       \
    st<C::Fp>(t, fe_add<C::Fp>(ld<C::Fp>(a), ld<C::Fp>(b))); }                                                  \
  extern "C" void he_##NAME##_fp_sub(const uint64_t* a, const uint64_t* b, uint64_t* t) {                       \
    st<C::Fp>(t, fe_sub<C::Fp>(ld<C::Fp>(a), ld<C::Fp>(b))); }                                                  \
  extern "C" void he_##NAME##_fp_neg(const uint64_t* a, uint64_t* t) { st<C::Fp>(t, fe_neg<C::Fp>(ld<C::Fp>(a))); } \
  extern "C" void he_##NAME##_fp_inv(const uint64_t* a, uint64_t* t) { st<C::Fp>(t, fe_inv<C::Fp>(ld<C::Fp>(a))); } \
  extern "C" void he_##NAME##_fp_inv_euclid(const uint64_t* a, uint64_t* t) { st<C::Fp>(t, fe_inv_euclid<C::Fp>(ld<C::Fp>(a))); } \
  extern "C" void he_##NAME##_fr_inv(const uint64_t* a, uint64_t* t) { st<C::Fr>(t, fe_inv<C::Fr>(ld<C::Fr>(a))); } \
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

// ---- the affine pre-reduction tree (aff_plan.cuh), walked exactly like kernels_aff.cuh does, one segment --------
// keys/vals: n sorted pairs (key 0 first); R levels; then the surviving records are added to their buckets the way
// k_accumulate + fix-up would.  Returns 0, or a negative code when a structural invariant is broken:
//   -1 a bucket was written twice by the tree, -2 a record touches a bucket the tree has written,
//   -3 the records are not sorted by key.
template <class P>
static int aff_tree_emul(long n, const uint32_t* keys, const uint32_t* vals, const uint64_t* pts, int R, uint32_t NB,
                         uint64_t* out_affine) {
  struct St { uint32_t hk, hr, tk, tr; };
  std::vector<Affine<P>> tmp;
  std::vector<char> tmp_inf;
  std::vector<Xyzz<P>> bucket(NB, xyzz_inf<P>());
  std::vector<char> written(NB, 0);
  int err = 0;
  auto load = [&](uint32_t ref, bool& inf) {
    Affine<P> p;
    if (ref & AFF_TEMP) { p = tmp[ref & AFF_IDX]; inf = tmp_inf[ref & AFF_IDX]; }
    else { p = ldaff<P>(pts + (size_t)(ref & AFF_IDX) * P::L); inf = affine_is_inf<P>(p); }
    if ((ref >> 31) && !inf) p.y = fe_neg<P>(p.y);
    return p;
  };
  auto to_bucket = [&](uint32_t key, const Affine<P>& p, bool inf) {
    if (written[key - 1]) err = -1;
    written[key - 1] = 1;
    bucket[key - 1] = inf ? xyzz_inf<P>() : xyzz_from_affine<P>(p);
  };
  std::vector<St> cur(n), nxt;
  for (long i = 0; i < n; i++) cur[i] = St{keys[i], vals[i], keys[i], vals[i]};
  for (int r = 0; r < R; r++) {
    size_t nin = cur.size(), nm = (nin + 1) / 2;
    nxt.assign(nm, St{0, 0, 0, 0});
    for (size_t i = 0; i < nm; i++) {
      St l = cur[2 * i], rr = 2 * i + 1 < nin ? cur[2 * i + 1] : St{0, 0, 0, 0};
      AffPlan pl = aff_plan(l.hk, l.hr, l.tk, l.tr, rr.hk, rr.hr, rr.tk, rr.tr);
      if (pl.add) {
        bool i1, i2;
        Affine<P> p1 = load(l.tr, i1), p2 = load(rr.hr, i2), sum = p1;
        Fe<P> d;
        int cls = aff_classify<P>(p1, i1, p2, i2, d);
        bool sinf = false;
        if (cls >= AFF_ADD) sum = aff_finish<P, false>(cls, p1, p2, fe_inv<P>(d));
        else if (cls == AFF_COPY2) sum = p2;
        else if (cls == AFF_INF) sinf = true;
        if (pl.sum_key) to_bucket(pl.sum_key, sum, sinf);
        uint32_t sref = AFF_TEMP | (uint32_t)tmp.size();
        tmp.push_back(sum);
        tmp_inf.push_back(sinf);
        if (pl.hr == AFF_SUM) pl.hr = sref;
        if (pl.tr == AFF_SUM) pl.tr = sref;
      } else {
        for (int k = 0; k < 2; k++)
          if (pl.st_key[k]) { bool inf; Affine<P> p = load(pl.st_ref[k], inf); to_bucket(pl.st_key[k], p, inf); }
      }
      nxt[i] = St{pl.hk, pl.hr, pl.tk, pl.tr};
    }
    cur.swap(nxt);
  }
  uint32_t last_key = 0;
  for (const St& b : cur) {
    uint32_t rk[2] = {b.hk == b.tk ? 0u : b.hk, b.tk}, rf[2] = {b.hr, b.tr};
    for (int k = 0; k < 2; k++) {
      if (!rk[k]) continue;
      if (rk[k] < last_key) err = -3;
      last_key = rk[k];
      if (written[rk[k] - 1]) err = -2;
      bool inf;
      Affine<P> p = load(rf[k], inf);
      if (!inf) xyzz_madd<P>(bucket[rk[k] - 1], p);
    }
  }
  for (uint32_t b = 0; b < NB; b++) staff<P>(out_affine + (size_t)b * P::L, bucket[b]);
  return err;
}
#define DEFINE_AFF(NAME, C)                                                                                        \
  extern "C" int he_##NAME##_aff_tree(long n, const uint32_t* keys, const uint32_t* vals, const uint64_t* pts, int R, \
                                      uint32_t NB, uint64_t* out_affine) {                                         \
    return aff_tree_emul<C::Fp>(n, keys, vals, pts, R, NB, out_affine); }
DEFINE_AFF(bn128, Bn254)
DEFINE_AFF(bls12_381, Bls12381)

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

// GLV split of one scalar (8 x 32-bit limbs in, any value < 2^256): out = |k1| (2 words), |k2| (2 words), flags[0..1] = signs;
// and phi(x) = beta * x on a Montgomery coordinate
#define DEFINE_GLV(NAME, C)                                                                                       \
  extern "C" void he_##NAME##_glv_split(const uint64_t* k, uint64_t* k1, uint64_t* k2, int* flags) {              \
    uint32_t kl[8], a[4], b[4];                                                                                   \
    memcpy(kl, k, 32);                                                                                            \
    bool n1, n2;                                                                                                  \
    glv_decompose<GlvOf<C>::type>(kl, a, n1, b, n2);                                                              \
    memcpy(k1, a, 16); memcpy(k2, b, 16);                                                                         \
    flags[0] = n1; flags[1] = n2; }                                                                               \
  extern "C" void he_##NAME##_glv_beta_x(const uint64_t* x, uint64_t* t) {                                        \
    st<C::Fp>(t, glv_beta_x<C::Fp, GlvOf<C>::type>(ld<C::Fp>(x))); }
DEFINE_GLV(bn128, Bn254)
DEFINE_GLV(bls12_381, Bls12381)

// Model of the low-latency bucket reduction (kernels_red.cuh K5') over the integers mod m: the same three steps with the
// same index functions (red_plan.cuh), group addition = addition mod m, doubling = times two.  buckets: 2^(c-1) values.
extern "C" uint64_t he_red2d_model(int c, int nch, const uint64_t* buckets, uint64_t m) {
  const RedPlan pl = red_plan(c);
  const uint32_t NR = 1u << pl.hr, NC = 1u << pl.hc;
  auto add = [&](uint64_t a, uint64_t b) { return (uint64_t)(((unsigned __int128)a + b) % m); };
  std::vector<uint64_t> RC(NR + NC, 0), seen(NR + NC, 0);
  for (unsigned blk = 0; blk < pl.row_blocks + pl.col_blocks; blk++) {
    uint64_t part[32];
    RedTask tk[32];
    for (int tq = 0; tq < 32; tq++) {
      tk[tq] = red_rowcol_task(pl, blk, tq);
      uint64_t acc = 0;
      for (uint32_t i = tk[tq].part; tk[tq].valid && i < tk[tq].entries; i += tk[tq].tpo) acc = add(acc, buckets[tk[tq].first + (size_t)i * tk[tq].stride]);
      part[tq] = acc;
    }
    for (int tq = 0; tq < 32; tq++) {          // red_block_tree: the group's first team ends up with the group's sum
      if (tk[tq].part != 0 || !tk[tq].valid) continue;
      uint64_t acc = 0;
      for (int k = 0; k < tk[tq].tpo; k++) acc = add(acc, part[tq + k]);
      const uint32_t slot = tk[tq].rows ? tk[tq].out : NR + tk[tq].out;
      RC[slot] = acc;
      seen[slot]++;
    }
  }
  for (uint32_t i = 0; i < NR + NC; i++) if (seen[i] != 1) return ~(uint64_t)0;   // every sum written exactly once
  std::vector<uint64_t> T(c, 0);
  for (int j = 0; j < c; j++) {
    const RedBits bt = red_bits_task(pl, j);
    for (uint32_t e = 0; e < bt.entries; e++) T[j] = add(T[j], RC[bt.base + red_bit_member(e, bt.bit)]);
  }
  const int nb = c - 1;
  std::vector<uint64_t> V(nch, 0);
  for (int q = 0; q < nch; q++) {
    int lo, hi;
    red_piece(nb, nch, q, lo, hi);
    uint64_t acc = 0;
    for (int b = hi - 1; b >= lo; b--) acc = add(add(acc, acc), T[b]);
    if (q == 0) acc = add(acc, T[c - 1]);
    V[q] = acc;
  }
  uint64_t acc = V[nch - 1];
  for (int r = nch - 2; r >= 0; r--) {
    int lo, hi;
    red_piece(nb, nch, r, lo, hi);
    for (int d = 0; d < hi - lo; d++) acc = add(acc, acc);
    acc = add(acc, V[r]);
  }
  return acc;
}
