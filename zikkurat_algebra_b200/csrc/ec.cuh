// Short-Weierstrass (a = 0) group law in extended Jacobian "XYZZ" coordinates
//   (X, Y, ZZ, ZZZ)  ~  affine (X/ZZ, Y/ZZZ),  ZZ^3 = ZZZ^2,  infinity <=> ZZ = 0.
//
// Device-side replacement for the reference's per-bucket arithmetic
//   bn128_G1_proj_madd_proj_aff   lib/cbits/curves/g1/proj/bn128_G1_proj.c:333-373  (11 Fp mul)
//   bn128_G1_proj_add             lib/cbits/curves/g1/proj/bn128_G1_proj.c:272-313
//   bn128_G1_proj_dbl             lib/cbits/curves/g1/proj/bn128_G1_proj.c:230-263
//   bn128_G1_jac_madd_jac_aff     lib/cbits/curves/g1/jac/bn128_G1_jac.c:362-422
// Formulas: EFD "xyzz" madd-2008-s (8M+2S), add-2008-s (12M+2S), dbl-2008-s-1, mdbl-2008-s-1; the two
// products of every Y3 share one Montgomery reduction (fe_mul2).
// The reference treats P+P, P+(-P), inf+P and P+inf explicitly in every addition; so do these
// routines (the formulas alone would return (0,0,0,0) for P+P).  Intermediate representatives differ
// from the reference's, which is fine: only canonical affine output is comparable.
#pragma once
#include "fp.cuh"

namespace zk {

template <class P>
struct Affine {
  Fe<P> x, y;
};

template <class P>
struct Xyzz {
  Fe<P> X, Y, ZZ, ZZZ;
};

template <class P>
ZK_HD bool xyzz_is_inf(const Xyzz<P>& a) {
  return fe_is_zero<P>(a.ZZ);
}

template <class P>
ZK_HD Xyzz<P> xyzz_inf() {
  Xyzz<P> r;
  r.X = fe_zero<P>();
  r.Y = fe_zero<P>();
  r.ZZ = fe_zero<P>();
  r.ZZZ = fe_zero<P>();
  return r;
}

template <class P>
ZK_HD Xyzz<P> xyzz_from_affine(const Affine<P>& p) {
  Xyzz<P> r;
  r.X = p.x;
  r.Y = p.y;
  r.ZZ = fe_one<P>();
  r.ZZZ = fe_one<P>();
  return r;
}

// The reference's affine infinity is "all bytes 0xFF" (bn128_G1_affine.c:43-49,62-65).  A canonical
// coordinate never has an all-ones top limb, so that limb is the cheap filter and the full test only
// runs for candidates.
template <class P>
ZK_HD bool affine_is_inf(const Affine<P>& p) {
  if (p.x.l[P::L - 1] != 0xffffffffu) return false;
  uint32_t a = 0xffffffffu;
#pragma unroll
  for (int i = 0; i < P::L; i++) a &= p.x.l[i] & p.y.l[i];
  return a == 0xffffffffu;
}

// 2*P for affine P (mdbl-2008-s-1, a = 0)
template <class P>
ZK_HD Xyzz<P> xyzz_dbl_affine(const Affine<P>& p) {
  Xyzz<P> r;
  Fe<P> U = fe_dbl<P>(p.y);
  Fe<P> V = fe_sqr<P>(U);
  Fe<P> W = fe_mul<P>(U, V);
  Fe<P> S = fe_mul<P>(p.x, V);
  Fe<P> XX = fe_sqr<P>(p.x);
  Fe<P> M = fe_add<P>(fe_dbl<P>(XX), XX);
  r.X = fe_sub<P>(fe_sub<P>(fe_sqr<P>(M), S), S);
  r.Y = fe_mul2<P>(M, fe_sub<P>(S, r.X), W, fe_neg<P>(p.y));   // M*(S - X3) - W*y, one reduction
  r.ZZ = V;
  r.ZZZ = W;
  return r;  // y = 0 (a 2-torsion point) gives ZZ = 0 = infinity, as it must
}

// 2*A (dbl-2008-s-1, a = 0)
template <class P>
ZK_HD Xyzz<P> xyzz_dbl(const Xyzz<P>& a) {
  if (xyzz_is_inf<P>(a)) return a;
  Xyzz<P> r;
  Fe<P> U = fe_dbl<P>(a.Y);
  Fe<P> V = fe_sqr<P>(U);
  Fe<P> W = fe_mul<P>(U, V);
  Fe<P> S = fe_mul<P>(a.X, V);
  Fe<P> XX = fe_sqr<P>(a.X);
  Fe<P> M = fe_add<P>(fe_dbl<P>(XX), XX);
  r.X = fe_sub<P>(fe_sub<P>(fe_sqr<P>(M), S), S);
  r.Y = fe_mul2<P>(M, fe_sub<P>(S, r.X), W, fe_neg<P>(a.Y));   // M*(S - X3) - W*Y1, one reduction
  r.ZZ = fe_mul<P>(V, a.ZZ);
  r.ZZZ = fe_mul<P>(W, a.ZZZ);
  return r;
}

// same operation with every multiplication out of line (fe_mul_call): small code, same arithmetic
template <class P>
ZK_HD void xyzz_madd_calls(Xyzz<P>& acc, const Affine<P>& p) {
  if (xyzz_is_inf<P>(acc)) {
    acc = xyzz_from_affine<P>(p);
    return;
  }
  Fe<P> Pd = fe_sub<P>(fe_mul_call<P>(p.x, acc.ZZ), acc.X);
  Fe<P> R = fe_sub<P>(fe_mul_call<P>(p.y, acc.ZZZ), acc.Y);
  if (fe_is_zero<P>(Pd)) {
    if (fe_is_zero<P>(R)) {  // same point: double (mdbl-2008-s-1)
      Fe<P> U = fe_dbl<P>(p.y);
      Fe<P> V = fe_sqr_call<P>(U);
      Fe<P> W = fe_mul_call<P>(U, V);
      Fe<P> S = fe_mul_call<P>(p.x, V);
      Fe<P> XX = fe_sqr_call<P>(p.x);
      Fe<P> M = fe_add<P>(fe_dbl<P>(XX), XX);
      acc.X = fe_sub<P>(fe_sub<P>(fe_sqr_call<P>(M), S), S);
      acc.Y = fe_mul2_call<P>(M, fe_sub<P>(S, acc.X), W, fe_neg<P>(p.y));
      acc.ZZ = V;
      acc.ZZZ = W;
    } else {
      acc = xyzz_inf<P>();
    }
    return;
  }
  Fe<P> PP = fe_sqr_call<P>(Pd);
  Fe<P> PPP = fe_mul_call<P>(Pd, PP);
  Fe<P> Q = fe_mul_call<P>(acc.X, PP);
  Fe<P> X3 = fe_sub<P>(fe_sub<P>(fe_sub<P>(fe_sqr_call<P>(R), PPP), Q), Q);
  Fe<P> Y3 = fe_mul2_call<P>(R, fe_sub<P>(Q, X3), fe_neg<P>(acc.Y), PPP);   // R*(Q - X3) - Y1*PPP, one reduction
  acc.ZZ = fe_mul_call<P>(acc.ZZ, PP);
  acc.ZZZ = fe_mul_call<P>(acc.ZZZ, PPP);
  acc.X = X3;
  acc.Y = Y3;
}

// acc += p   (p affine, already known not to be infinity).  The bucket-insertion primitive.
template <class P, bool CALLS = false>
ZK_HD void xyzz_madd(Xyzz<P>& acc, const Affine<P>& p) {
  if (CALLS) { xyzz_madd_calls<P>(acc, p); return; }
  if (xyzz_is_inf<P>(acc)) {
    acc = xyzz_from_affine<P>(p);
    return;
  }
  Fe<P> Pd = fe_sub<P>(fe_mul<P>(p.x, acc.ZZ), acc.X);   // U2 - X1
  Fe<P> R = fe_sub<P>(fe_mul<P>(p.y, acc.ZZZ), acc.Y);   // S2 - Y1
  if (fe_is_zero<P>(Pd)) {
    if (fe_is_zero<P>(R)) acc = xyzz_dbl_affine<P>(p);   // same point: double
    else acc = xyzz_inf<P>();                            // opposite points
    return;
  }
  Fe<P> PP = fe_sqr<P>(Pd);
  Fe<P> PPP = fe_mul<P>(Pd, PP);
  Fe<P> Q = fe_mul<P>(acc.X, PP);
  Fe<P> X3 = fe_sub<P>(fe_sub<P>(fe_sub<P>(fe_sqr<P>(R), PPP), Q), Q);
  Fe<P> Y3 = fe_mul2<P>(R, fe_sub<P>(Q, X3), fe_neg<P>(acc.Y), PPP);   // R*(Q - X3) - Y1*PPP, one reduction
  acc.ZZ = fe_mul<P>(acc.ZZ, PP);
  acc.ZZZ = fe_mul<P>(acc.ZZZ, PPP);
  acc.X = X3;
  acc.Y = Y3;
}

// a + b, both XYZZ (add-2008-s)
template <class P>
ZK_HD Xyzz<P> xyzz_add(const Xyzz<P>& a, const Xyzz<P>& b) {
  if (xyzz_is_inf<P>(a)) return b;
  if (xyzz_is_inf<P>(b)) return a;
  Fe<P> U1 = fe_mul<P>(a.X, b.ZZ);
  Fe<P> U2 = fe_mul<P>(b.X, a.ZZ);
  Fe<P> S1 = fe_mul<P>(a.Y, b.ZZZ);
  Fe<P> S2 = fe_mul<P>(b.Y, a.ZZZ);
  Fe<P> Pd = fe_sub<P>(U2, U1);
  Fe<P> R = fe_sub<P>(S2, S1);
  if (fe_is_zero<P>(Pd)) {
    if (fe_is_zero<P>(R)) return xyzz_dbl<P>(a);
    return xyzz_inf<P>();
  }
  Xyzz<P> r;
  Fe<P> PP = fe_sqr<P>(Pd);
  Fe<P> PPP = fe_mul<P>(Pd, PP);
  Fe<P> Q = fe_mul<P>(U1, PP);
  r.X = fe_sub<P>(fe_sub<P>(fe_sub<P>(fe_sqr<P>(R), PPP), Q), Q);
  r.Y = fe_mul2<P>(R, fe_sub<P>(Q, r.X), fe_neg<P>(S1), PPP);   // R*(Q - X3) - S1*PPP, one reduction
  r.ZZ = fe_mul<P>(fe_mul<P>(a.ZZ, b.ZZ), PP);
  r.ZZZ = fe_mul<P>(fe_mul<P>(a.ZZZ, b.ZZZ), PPP);
  return r;
}

// add-2008-s with every multiplication out of line (same arithmetic, small code / fewer registers)
template <class P>
ZK_HD Xyzz<P> xyzz_add_calls(const Xyzz<P>& a, const Xyzz<P>& b) {
  if (xyzz_is_inf<P>(a)) return b;
  if (xyzz_is_inf<P>(b)) return a;
  Fe<P> U1 = fe_mul_call<P>(a.X, b.ZZ);
  Fe<P> U2 = fe_mul_call<P>(b.X, a.ZZ);
  Fe<P> S1 = fe_mul_call<P>(a.Y, b.ZZZ);
  Fe<P> S2 = fe_mul_call<P>(b.Y, a.ZZZ);
  Fe<P> Pd = fe_sub<P>(U2, U1);
  Fe<P> R = fe_sub<P>(S2, S1);
  if (fe_is_zero<P>(Pd)) {
    if (fe_is_zero<P>(R)) return xyzz_dbl<P>(a);
    return xyzz_inf<P>();
  }
  Xyzz<P> r;
  Fe<P> PP = fe_sqr_call<P>(Pd);
  Fe<P> PPP = fe_mul_call<P>(Pd, PP);
  Fe<P> Q = fe_mul_call<P>(U1, PP);
  r.X = fe_sub<P>(fe_sub<P>(fe_sub<P>(fe_sqr_call<P>(R), PPP), Q), Q);
  r.Y = fe_mul2_call<P>(R, fe_sub<P>(Q, r.X), fe_neg<P>(S1), PPP);
  r.ZZ = fe_mul_call<P>(fe_mul_call<P>(a.ZZ, b.ZZ), PP);
  r.ZZZ = fe_mul_call<P>(fe_mul_call<P>(a.ZZZ, b.ZZZ), PPP);
  return r;
}

template <class P>
ZK_HD Affine<P> affine_neg(const Affine<P>& p) {
  Affine<P> r;
  r.x = p.x;
  r.y = fe_neg<P>(p.y);
  return r;
}

// ---- conversions to the reference's output representations ------------
// homogeneous projective (X:Y:Z), infinity = (0, R, 0)           bn128_G1_proj.c:178-182
template <class P>
ZK_HD void xyzz_to_proj(const Xyzz<P>& a, Fe<P>& X, Fe<P>& Y, Fe<P>& Z) {
  if (xyzz_is_inf<P>(a)) {
    X = fe_zero<P>(); Y = fe_one<P>(); Z = fe_zero<P>();
    return;
  }
  X = fe_mul<P>(a.X, a.ZZZ);
  Y = fe_mul<P>(a.Y, a.ZZ);
  Z = fe_mul<P>(a.ZZ, a.ZZZ);
}
// Jacobian (X:Y:Z) ~ (X/Z^2, Y/Z^3), infinity = (R, R, 0)         bn128_G1_jac.c:183-187
// With ZZ = z^2, ZZZ = z^3 the representative (X*ZZ, Y*ZZZ, ZZ) has Z' = z^2.
template <class P>
ZK_HD void xyzz_to_jac(const Xyzz<P>& a, Fe<P>& X, Fe<P>& Y, Fe<P>& Z) {
  if (xyzz_is_inf<P>(a)) {
    X = fe_one<P>(); Y = fe_one<P>(); Z = fe_zero<P>();
    return;
  }
  X = fe_mul<P>(a.X, a.ZZ);
  Y = fe_mul<P>(a.Y, a.ZZZ);
  Z = a.ZZ;
}
// canonical affine; returns false for infinity (caller writes the 0xFF pattern)  bn128_G1_proj.c:132-144
template <class P>
ZK_HD bool xyzz_to_affine(const Xyzz<P>& a, Affine<P>& out) {
  if (xyzz_is_inf<P>(a)) return false;
  // 1/ZZ = ZZ^-3 * ZZ^2 ... one inversion of ZZZ*ZZ gives both: (ZZ*ZZZ)^-1 * ZZZ = 1/ZZ, * ZZ = 1/ZZZ
  Fe<P> inv = fe_inv<P>(fe_mul<P>(a.ZZ, a.ZZZ));
  out.x = fe_mul<P>(a.X, fe_mul<P>(inv, a.ZZZ));
  out.y = fe_mul<P>(a.Y, fe_mul<P>(inv, a.ZZ));
  return true;
}
// reference projective / Jacobian input -> XYZZ (used by the partial-sum combine entry point)
template <class P>
ZK_HD Xyzz<P> xyzz_from_proj(const Fe<P>& X, const Fe<P>& Y, const Fe<P>& Z) {
  if (fe_is_zero<P>(Z)) return xyzz_inf<P>();
  // x = X/Z, y = Y/Z: choose z = Z  =>  (X*Z, Y*Z^2, Z^2, Z^3)
  Xyzz<P> r;
  Fe<P> Z2 = fe_sqr<P>(Z);
  r.X = fe_mul<P>(X, Z);
  r.Y = fe_mul<P>(Y, Z2);
  r.ZZ = Z2;
  r.ZZZ = fe_mul<P>(Z2, Z);
  return r;
}
template <class P>
ZK_HD Xyzz<P> xyzz_from_jac(const Fe<P>& X, const Fe<P>& Y, const Fe<P>& Z) {
  if (fe_is_zero<P>(Z)) return xyzz_inf<P>();
  Xyzz<P> r;
  r.X = X;
  r.Y = Y;
  r.ZZ = fe_sqr<P>(Z);
  r.ZZZ = fe_mul<P>(r.ZZ, Z);
  return r;
}

}  // namespace zk
