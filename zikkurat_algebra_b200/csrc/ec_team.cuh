// 4-lane "team" versions of the XYZZ group operations for the latency-bound phases of the MSM
// (head fix-up tree, upper bucket-reduction levels, window combination).
//
// A single thread needs ~14 dependent field multiplications for one addition; a warp instruction costs
// the same IMAD-pipe time whether 1 or 32 lanes are active, so those phases were limited by the
// latency of one thread's multiplication chain.  Here the 4 lanes of a team hold IDENTICAL copies of
// the operands; in every round each lane multiplies a different pair of values, the four products are
// exchanged with warp shuffles, and the cheap additions/subtractions are done redundantly by all lanes.
//   add : 14 multiplications -> 4 rounds      (add-2008-s)
//   dbl :  9 multiplications -> 3 rounds      (dbl-2008-s-1, a = 0)
// Exceptional cases (infinity operands, P+P, P+(-P)) are decided on identical data, so a team never
// diverges internally; different teams of a warp may diverge, hence the team-wide shuffle mask.
// Same group law as ec.cuh (reference: lib/cbits/curves/g1/proj/bn128_G1_proj.c:230-313).
#pragma once
#include "ec.cuh"

namespace zk {

struct Team {
  unsigned mask;  // the 4 lanes of this team
  int base;       // lane id of team member 0
  int t;          // my index in the team (0..3)
  __device__ __forceinline__ Team() {
    int lane = threadIdx.x & 31;
    t = lane & 3;
    base = lane & ~3;
    mask = 0xFu << base;
  }
};

template <class P>
ZK_D Fe<P> team_get(const Team& tm, const Fe<P>& v, int member) {
  Fe<P> r;
#pragma unroll
  for (int i = 0; i < P::L; i++) r.l[i] = __shfl_sync(tm.mask, v.l[i], tm.base + member);
  return r;
}
template <class P>
ZK_D Fe<P> team_sel(int t, const Fe<P>& a, const Fe<P>& b, const Fe<P>& c, const Fe<P>& d) {
  Fe<P> r;
#pragma unroll
  for (int i = 0; i < P::L; i++) {
    uint32_t lo = t == 0 ? a.l[i] : b.l[i];
    uint32_t hi = t == 2 ? c.l[i] : d.l[i];
    r.l[i] = t < 2 ? lo : hi;
  }
  return r;
}

template <class P>
ZK_D Xyzz<P> xyzz_dbl_team(const Team& tm, const Xyzz<P>& a) {
  if (xyzz_is_inf<P>(a)) return a;
  Xyzz<P> r;
  Fe<P> U = fe_dbl<P>(a.Y);
  // round 1: V = U^2, XX = X^2
  Fe<P> x = team_sel<P>(tm.t, U, a.X, U, a.X);
  Fe<P> pr = fe_sqr<P>(x);
  Fe<P> V = team_get<P>(tm, pr, 0), XX = team_get<P>(tm, pr, 1);
  Fe<P> M = fe_add<P>(fe_dbl<P>(XX), XX);
  // round 2: W = U*V, S = X*V, MM = M^2, ZZ3 = V*ZZ
  x = team_sel<P>(tm.t, U, a.X, M, V);
  Fe<P> y = team_sel<P>(tm.t, V, V, M, a.ZZ);
  pr = fe_mul<P>(x, y);
  Fe<P> W = team_get<P>(tm, pr, 0), S = team_get<P>(tm, pr, 1), MM = team_get<P>(tm, pr, 2);
  r.ZZ = team_get<P>(tm, pr, 3);
  r.X = fe_sub<P>(fe_sub<P>(MM, S), S);
  // round 3: M*(S - X3), W*Y, W*ZZZ
  x = team_sel<P>(tm.t, M, W, W, W);
  y = team_sel<P>(tm.t, fe_sub<P>(S, r.X), a.Y, a.ZZZ, a.ZZZ);
  pr = fe_mul<P>(x, y);
  r.Y = fe_sub<P>(team_get<P>(tm, pr, 0), team_get<P>(tm, pr, 1));
  r.ZZZ = team_get<P>(tm, pr, 2);
  return r;
}

template <class P>
ZK_D Xyzz<P> xyzz_add_team(const Team& tm, const Xyzz<P>& a, const Xyzz<P>& b) {
  if (xyzz_is_inf<P>(a)) return b;
  if (xyzz_is_inf<P>(b)) return a;
  // round 1: U1 = X1*ZZ2, U2 = X2*ZZ1, S1 = Y1*ZZZ2, S2 = Y2*ZZZ1
  Fe<P> x = team_sel<P>(tm.t, a.X, b.X, a.Y, b.Y);
  Fe<P> y = team_sel<P>(tm.t, b.ZZ, a.ZZ, b.ZZZ, a.ZZZ);
  Fe<P> pr = fe_mul<P>(x, y);
  Fe<P> U1 = team_get<P>(tm, pr, 0), U2 = team_get<P>(tm, pr, 1), S1 = team_get<P>(tm, pr, 2), S2 = team_get<P>(tm, pr, 3);
  Fe<P> Pd = fe_sub<P>(U2, U1), R = fe_sub<P>(S2, S1);
  if (fe_is_zero<P>(Pd)) {
    if (fe_is_zero<P>(R)) return xyzz_dbl_team<P>(tm, a);
    return xyzz_inf<P>();
  }
  Xyzz<P> r;
  // round 2: PP = P^2, RR = R^2, ZZ1*ZZ2, ZZZ1*ZZZ2
  x = team_sel<P>(tm.t, Pd, R, a.ZZ, a.ZZZ);
  y = team_sel<P>(tm.t, Pd, R, b.ZZ, b.ZZZ);
  pr = fe_mul<P>(x, y);
  Fe<P> PP = team_get<P>(tm, pr, 0), RR = team_get<P>(tm, pr, 1), Z2 = team_get<P>(tm, pr, 2), Z3 = team_get<P>(tm, pr, 3);
  // round 3: PPP = P*PP, Q = U1*PP, ZZ3 = ZZ1*ZZ2*PP
  x = team_sel<P>(tm.t, Pd, U1, Z2, Z2);
  pr = fe_mul<P>(x, PP);
  Fe<P> PPP = team_get<P>(tm, pr, 0), Q = team_get<P>(tm, pr, 1);
  r.ZZ = team_get<P>(tm, pr, 2);
  r.X = fe_sub<P>(fe_sub<P>(fe_sub<P>(RR, PPP), Q), Q);
  // round 4: R*(Q - X3), S1*PPP, ZZZ3 = ZZZ1*ZZZ2*PPP
  x = team_sel<P>(tm.t, R, S1, Z3, Z3);
  y = team_sel<P>(tm.t, fe_sub<P>(Q, r.X), PPP, PPP, PPP);
  pr = fe_mul<P>(x, y);
  r.Y = fe_sub<P>(team_get<P>(tm, pr, 0), team_get<P>(tm, pr, 1));
  r.ZZZ = team_get<P>(tm, pr, 2);
  return r;
}

}  // namespace zk
