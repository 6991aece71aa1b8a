// Bookkeeping of the affine pre-reduction tree (kernels_aff.cuh): pure integer logic, shared by the device
// kernels and by the host-side emulation test (tests/host_emul).
//
// The sorted pairs of a segment are reduced pairwise, level by level.  A *block* of level r covers 2^r
// consecutive sorted pairs and is summarised by the two runs (maximal groups of equal keys) that may
// continue outside of it:
//     head = (key, ref) of its first run,   tail = (key, ref) of its last run;   head key == tail key <=> one run only.
// `ref` names the affine point holding the partial sum of that run inside the block: an index into the
// caller's point array (bit 31 = negate, as in the sorted pairs) or, with bit 30 set, into the tree's
// temporary array.  Runs that lie strictly inside a block are complete: the only writer of their bucket
// is the merge that closed them.  Merging two neighbouring blocks needs AT MOST ONE point addition
// (left tail + right head, when the keys agree), which is what makes every level one uniform batch of
// independent affine additions sharing one field inversion.
// Key 0 is the "no insertion" digit (and the padding behind the end of a segment): never added, never stored.
#pragma once
#include <stdint.h>

#include "ec.cuh"

namespace zk {

constexpr uint32_t AFF_TEMP = 0x40000000u;  // ref lives in the temporary array
constexpr uint32_t AFF_IDX = 0x3fffffffu;
constexpr uint32_t AFF_SUM = 0xffffffffu;   // placeholder: "the sum computed by this merge" (never a valid ref)

struct AffPlan {
  uint32_t hk, hr, tk, tr;        // merged block; a ref equal to AFF_SUM stands for the new sum
  uint32_t sum_key;               // != 0: the new sum is a complete run, it goes to bucket sum_key
  uint32_t st_key[2], st_ref[2];  // complete runs to copy to their buckets as they are (key 0 = none)
  bool add;                       // the sum  point(Ltr) + point(Rhr)  is needed
};

ZK_HD AffPlan aff_plan(uint32_t Lhk, uint32_t Lhr, uint32_t Ltk, uint32_t Ltr, uint32_t Rhk, uint32_t Rhr, uint32_t Rtk,
                       uint32_t Rtr) {
  AffPlan p;
  const bool Ls = Lhk == Ltk, Rs = Rhk == Rtk;
  p.add = Ltk == Rhk && Ltk != 0;
  p.sum_key = 0;
  p.st_key[0] = p.st_key[1] = 0;
  p.st_ref[0] = p.st_ref[1] = 0;
  p.hk = Lhk; p.hr = Lhr; p.tk = Rtk; p.tr = Rtr;
  if (p.add) {
    if (Ls) p.hr = AFF_SUM;
    if (Rs) p.tr = AFF_SUM;
    if (Ls && Rs) p.tk = Lhk;
    if (!Ls && !Rs) p.sum_key = Ltk;
  } else {
    if (!Ls && Ltk != 0) { p.st_key[0] = Ltk; p.st_ref[0] = Ltr; }
    if (!Rs && Rhk != 0) { p.st_key[1] = Rhk; p.st_ref[1] = Rhr; }
  }
  return p;
}

template <class P, bool CALLS>
ZK_HD Fe<P> aff_mul(const Fe<P>& a, const Fe<P>& b) {
  if constexpr (CALLS) return fe_mul_call<P>(a, b);
  else return fe_mul<P>(a, b);
}
// a*b and a*c at once (fp.cuh: mont_mul_pair_limbs) for prime fields with out-of-line multiplications; two plain products otherwise
template <class P, class = void> struct AffIsExt { static constexpr bool value = false; };
template <class P> struct AffIsExt<P, decltype((void)sizeof(typename P::Base))> { static constexpr bool value = true; };
template <class P, bool CALLS>
ZK_HD void aff_mul_pair(const Fe<P>& a, const Fe<P>& b, const Fe<P>& c, Fe<P>& ab, Fe<P>& ac) {
#ifdef ZK_AFF_NO_PAIR   // build-time experiment switch (tools/build_variant.py)
  constexpr bool PAIR = false;
#else
  constexpr bool PAIR = CALLS && !AffIsExt<P>::value;
#endif
  if constexpr (PAIR) {
    FePair<P> r = fe_mul_pair_call<P>(a, b, c);
    ab = r.u;
    ac = r.v;
  } else {
    const Fe<P> a0 = a;
    ab = aff_mul<P, CALLS>(a0, b);
    ac = aff_mul<P, CALLS>(a0, c);
  }
}
template <class P, bool CALLS>
ZK_HD Fe<P> aff_sqr(const Fe<P>& a) {
  if constexpr (CALLS) return fe_sqr_call<P>(a);
  else return fe_sqr<P>(a);
}

// ---- the per-merge arithmetic (host+device: unit-tested on the CPU through tests/host_emul) ----------------
enum AffClass : int { AFF_INF = 0, AFF_COPY1 = 1, AFF_COPY2 = 2, AFF_ADD = 3, AFF_DBL = 4 };

// class of P1 + P2 and, for AFF_ADD / AFF_DBL, the denominator of the slope (never zero)
template <class P>
ZK_HD int aff_classify(const Affine<P>& p1, bool inf1, const Affine<P>& p2, bool inf2, Fe<P>& d) {
  if (inf1) return inf2 ? AFF_INF : AFF_COPY2;
  if (inf2) return AFF_COPY1;
  if (fe_eq<P>(p1.x, p2.x)) {
    if (fe_eq<P>(p1.y, p2.y) && !fe_is_zero<P>(p1.y)) { d = fe_dbl<P>(p1.y); return AFF_DBL; }
    return AFF_INF;   // opposite points (or a 2-torsion point doubled)
  }
  d = fe_sub<P>(p2.x, p1.x);
  return AFF_ADD;
}
// the sum for AFF_ADD / AFF_DBL given dinv = 1/d:  lambda = (y2-y1)/(x2-x1) or 3*x1^2/(2*y1)
template <class P, bool CALLS>
ZK_HD Affine<P> aff_finish(int cls, const Affine<P>& p1, const Affine<P>& p2, const Fe<P>& dinv) {
  Fe<P> num;
  if (cls == AFF_DBL) {
    Fe<P> xx = aff_sqr<P, CALLS>(p1.x);
    num = fe_add<P>(fe_dbl<P>(xx), xx);
  } else {
    num = fe_sub<P>(p2.y, p1.y);
  }
  Fe<P> lam = aff_mul<P, CALLS>(num, dinv);
  Affine<P> r;
  r.x = fe_sub<P>(fe_sub<P>(aff_sqr<P, CALLS>(lam), p1.x), p2.x);
  r.y = fe_sub<P>(aff_mul<P, CALLS>(lam, fe_sub<P>(p1.x, r.x)), p1.y);
  return r;
}

}  // namespace zk
