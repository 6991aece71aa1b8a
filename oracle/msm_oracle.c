// =====================================================================================
// TEST INFRASTRUCTURE ONLY -- CPU oracle ("port") for the G1 MSM hot path.
//
// A plain-C restatement of the algorithm of bkomuves/zikkurat-algebra's generated C
// library for G1 multi-scalar multiplication on BN254 ("bn128") and BLS12-381:
// Montgomery Fp arithmetic on 64-bit limbs, homogeneous-projective and Jacobian group
// laws with the reference's exceptional-case handling, unsigned-window Pippenger with the
// reference's window heuristic, and conversion to canonical affine Montgomery bytes.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg
// may load this file's shared object; the product (zikkurat_algebra_b200/csrc) never does.
//
// Parity pinning: this restatement is checked (tests/test_oracle.py) against
//   (1) oracle/_ref/libzk_ref.so  -- the UNMODIFIED reference sources + platform.h shim,
//   (2) tests/pyec.py             -- an independent Python big-int affine implementation,
//   (3) tests/golden/*.json       -- vectors generated from (1) by tests/golden/make_golden.py.
// The reference has no golden vectors of its own for this path.
//
// All exported symbols are prefixed zko_ and take the reference's memory layout:
// little-endian uint64 limbs, Fp in Montgomery form, affine infinity = all bytes 0xFF.
// One code body serves both curves (limb count is a run-time field of `curve_t`), the way
// the reference generates its two twins from one template.
// =====================================================================================
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <pthread.h>

typedef unsigned __int128 u128;
#define MAXL 6

typedef struct {
  int L;               // 64-bit limbs per Fp element (4 | 6)
  uint64_t p[MAXL];    // base-field prime
  uint64_t pinv;       // -p^{-1} mod 2^64
  uint64_t one[MAXL];  // R mod p
  uint64_t r2[MAXL];   // R^2 mod p
  int b3;              // 3*B as a small integer (9 | 12)
  uint64_t r[4];       // scalar-field prime
  uint64_t rinv;       // -r^{-1} mod 2^64
} curve_t;

// constants: bn128_Fp_mont.c:20,130-131,145 ; bn128_Fr_mont.c:20,145 ; bn128_G1_proj.c:66-72
static const curve_t BN = {
  4,
  {0x3c208c16d87cfd47ULL, 0x97816a916871ca8dULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL},
  0x87d20782e4866389ULL,
  {0xd35d438dc58f0d9dULL, 0x0a78eb28f5c70b3dULL, 0x666ea36f7879462cULL, 0x0e0a77c19a07df2fULL},
  {0xf32cfc5b538afa89ULL, 0xb5e71911d44501fbULL, 0x47ab1eff0a417ff6ULL, 0x06d89f71cab8351fULL},
  9,
  {0x43e1f593f0000001ULL, 0x2833e84879b97091ULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL},
  0xc2e1f593efffffffULL,
};
// constants: bls12_381_Fp_mont.c:20,136-137,151 ; bls12_381_Fr_mont.c:20,145 ; bls12_381_G1_proj.c:66-73
static const curve_t BLS = {
  6,
  {0xb9feffffffffaaabULL, 0x1eabfffeb153ffffULL, 0x6730d2a0f6b0f624ULL, 0x64774b84f38512bfULL,
   0x4b1ba7b6434bacd7ULL, 0x1a0111ea397fe69aULL},
  0x89f3fffcfffcfffdULL,
  {0x760900000002fffdULL, 0xebf4000bc40c0002ULL, 0x5f48985753c758baULL, 0x77ce585370525745ULL,
   0x5c071a97a256ec6dULL, 0x15f65ec3fa80e493ULL},
  {0xf4df1f341c341746ULL, 0x0a76e6a609d104f1ULL, 0x8de5476c4c95b6d5ULL, 0x67eb88a9939d83c0ULL,
   0x9a793e85b519952dULL, 0x11988fe592cae3aaULL},
  12,
  {0xffffffff00000001ULL, 0x53bda402fffe5bfeULL, 0x3339d80809a1d805ULL, 0x73eda753299d7d48ULL},
  0xfffffffeffffffffULL,
};

// ------------------------------------------------------------------------------------
// multi-limb helpers (bigint256.c / bigint384.c add, sub with carry)

static inline uint64_t big_add(int L, const uint64_t *a, const uint64_t *b, uint64_t *t) {
  u128 c = 0;
  for (int i = 0; i < L; i++) { c += (u128)a[i] + b[i]; t[i] = (uint64_t)c; c >>= 64; }
  return (uint64_t)c;
}
static inline uint64_t big_sub(int L, const uint64_t *a, const uint64_t *b, uint64_t *t) {
  uint64_t bw = 0;
  for (int i = 0; i < L; i++) {
    u128 d = (u128)a[i] - b[i] - bw; t[i] = (uint64_t)d; bw = (uint64_t)(d >> 64) & 1;
  }
  return bw;
}
static inline int big_geq(int L, const uint64_t *a, const uint64_t *b) {
  for (int i = L - 1; i >= 0; i--) { if (a[i] > b[i]) return 1; if (a[i] < b[i]) return 0; }
  return 1;
}
static inline int big_is_zero(int L, const uint64_t *a) {
  uint64_t o = 0; for (int i = 0; i < L; i++) o |= a[i]; return o == 0;
}
static inline int big_eq(int L, const uint64_t *a, const uint64_t *b) {
  uint64_t o = 0; for (int i = 0; i < L; i++) o |= a[i] ^ b[i]; return o == 0;
}

// ------------------------------------------------------------------------------------
// Fp in Montgomery form

// bn128_Fp_mont.c:84-90 (add then sub_prime_if_above :72-81)
static void fp_add(const curve_t *C, const uint64_t *a, const uint64_t *b, uint64_t *t) {
  big_add(C->L, a, b, t);           // p < 2^(64L-1): never carries out
  if (big_geq(C->L, t, C->p)) big_sub(C->L, t, C->p, t);
}
// bn128_Fp_mont.c:98-104 (subtract, add the prime back on borrow)
static void fp_sub(const curve_t *C, const uint64_t *a, const uint64_t *b, uint64_t *t) {
  if (big_sub(C->L, a, b, t)) big_add(C->L, t, C->p, t);
}
// bn128_Fp_mont.c:44-60 (0 -> 0, else p - x)
static void fp_neg(const curve_t *C, const uint64_t *a, uint64_t *t) {
  if (big_is_zero(C->L, a)) memset(t, 0, 8 * C->L); else big_sub(C->L, C->p, a, t);
}
// word-serial Montgomery reduction of a 2L-limb value < R*p: bn128_Fp_mont.c:140-169
static void redc(int L, const uint64_t *mod, uint64_t minv, uint64_t *T /*2L+1*/, uint64_t *t) {
  T[2 * L] = 0;
  for (int i = 0; i < L; i++) {
    uint64_t m = T[i] * minv;
    u128 c = 0;
    for (int j = 0; j < L; j++) {
      c += (u128)m * mod[j] + T[i + j]; T[i + j] = (uint64_t)c; c >>= 64;
    }
    for (int j = i + L; c && j <= 2 * L; j++) { c += T[j]; T[j] = (uint64_t)c; c >>= 64; }
  }
  memcpy(t, T + L, 8 * L);
  if (T[2 * L] || big_geq(L, t, mod)) big_sub(L, t, mod, t);
}
// schoolbook product (bigint256.c:267-356) followed by REDC: bn128_Fp_mont.c:189-193
static void fp_mul(const curve_t *C, const uint64_t *a, const uint64_t *b, uint64_t *t) {
  uint64_t T[2 * MAXL + 1];
  int L = C->L;
  memset(T, 0, sizeof(T));
  for (int i = 0; i < L; i++) {
    u128 c = 0;
    for (int j = 0; j < L; j++) {
      c += (u128)a[i] * b[j] + T[i + j]; T[i + j] = (uint64_t)c; c >>= 64;
    }
    T[i + L] = (uint64_t)c;
  }
  redc(L, C->p, C->pinv, T, t);
}
static void fp_sqr(const curve_t *C, const uint64_t *a, uint64_t *t) { fp_mul(C, a, a, t); }

// Inverse in Montgomery form.  The reference goes through the standard representation with a
// binary-Euclid inverse and multiplies by R^3 (bn128_Fp_mont.c:201-204, bn128_Fp_std.c:252-315);
// the value is unique, so this restatement uses Fermat a^(p-2) instead (same result, no tables).
static void fp_inv(const curve_t *C, const uint64_t *a, uint64_t *t) {
  int L = C->L;
  uint64_t e[MAXL], acc[MAXL], base[MAXL];
  uint64_t two[MAXL] = {2, 0, 0, 0, 0, 0};
  big_sub(L, C->p, two, e);
  memcpy(acc, C->one, 8 * L);
  memcpy(base, a, 8 * L);
  for (int i = 0; i < 64 * L; i++) {
    if ((e[i >> 6] >> (i & 63)) & 1) fp_mul(C, acc, base, acc);
    fp_sqr(C, base, base);
  }
  memcpy(t, acc, 8 * L);
}

// Fr Montgomery -> standard: one REDC of (k, 0): bn128_Fr_mont.c:330-335
static void fr_to_std(const curve_t *C, const uint64_t *a, uint64_t *t) {
  uint64_t T[9];
  memcpy(T, a, 32); memset(T + 4, 0, 40);
  redc(4, C->r, C->rinv, T, t);
}

// ------------------------------------------------------------------------------------
// affine encoding of infinity: all bytes 0xFF (bn128_G1_affine.c:43-49,62-65,88-91)
static int aff_is_inf(const curve_t *C, const uint64_t *a) {
  uint64_t o = ~(uint64_t)0;
  for (int i = 0; i < 2 * C->L; i++) o &= a[i];
  return o == ~(uint64_t)0;
}

// ------------------------------------------------------------------------------------
// homogeneous projective coordinates (X:Y:Z), infinity = (0, R, 0): bn128_G1_proj.c:172-182
#define PX(P) (P)
#define PY(P) ((P) + L)
#define PZ(P) ((P) + 2 * L)

static void proj_set_inf(const curve_t *C, uint64_t *P) {
  int L = C->L; memset(P, 0, 24 * L); memcpy(PY(P), C->one, 8 * L);
}
static int proj_is_inf(const curve_t *C, const uint64_t *P) {
  int L = C->L;
  return big_is_zero(L, PZ(P)) && !big_is_zero(L, PY(P)) && big_is_zero(L, PX(P));
}
// bn128_G1_proj.c:120-128
static void proj_from_affine(const curve_t *C, const uint64_t *A, uint64_t *P) {
  int L = C->L;
  if (aff_is_inf(C, A)) { proj_set_inf(C, P); return; }
  memcpy(P, A, 16 * L); memcpy(PZ(P), C->one, 8 * L);
}
// bn128_G1_proj.c:132-144
static void proj_to_affine(const curve_t *C, const uint64_t *P, uint64_t *A) {
  int L = C->L;
  if (big_is_zero(L, PZ(P))) { memset(A, 0xff, 16 * L); return; }
  uint64_t zi[MAXL];
  fp_inv(C, PZ(P), zi);
  fp_mul(C, PX(P), zi, A);
  fp_mul(C, PY(P), zi, A + L);
}
static void fp_scale_small(const curve_t *C, const uint64_t *a, int k, uint64_t *t) {
  // k*a by repeated addition (reference: scale_by_3B, bn128_G1_proj.c:66-72)
  uint64_t acc[MAXL], base[MAXL];
  int L = C->L, first = 1;
  memcpy(base, a, 8 * L);
  memset(acc, 0, 8 * L);
  while (k) {
    if (k & 1) { if (first) { memcpy(acc, base, 8 * L); first = 0; } else fp_add(C, acc, base, acc); }
    fp_add(C, base, base, base);
    k >>= 1;
  }
  memcpy(t, acc, 8 * L);
}
// doubling, EFD dbl-2007-bl with a = 0: bn128_G1_proj.c:230-263
static void proj_dbl(const curve_t *C, const uint64_t *P, uint64_t *Q) {
  int L = C->L;
  uint64_t XX[MAXL], w[MAXL], s[MAXL], ss[MAXL], sss[MAXL], Rr[MAXL], RR[MAXL], B[MAXL], h[MAXL], t[MAXL];
  fp_sqr(C, PX(P), XX);
  fp_add(C, XX, XX, w); fp_add(C, w, XX, w);          // w = 3 XX   (a = 0)
  fp_mul(C, PY(P), PZ(P), s); fp_add(C, s, s, s);     // s = 2 Y Z
  fp_sqr(C, s, ss);
  fp_mul(C, s, ss, sss);
  fp_mul(C, PY(P), s, Rr);
  fp_sqr(C, Rr, RR);
  fp_add(C, PX(P), Rr, B); fp_sqr(C, B, B); fp_sub(C, B, XX, B); fp_sub(C, B, RR, B);
  fp_sqr(C, w, h); fp_sub(C, h, B, h); fp_sub(C, h, B, h);
  fp_sub(C, B, h, t); fp_mul(C, t, w, t); fp_sub(C, t, RR, t); fp_sub(C, t, RR, t);
  fp_mul(C, h, s, PX(Q));
  memcpy(PY(Q), t, 8 * L);
  memcpy(PZ(Q), sss, 8 * L);
}
// complete addition for a = 0, Renes-Costello-Batina 2015 algorithm 7: bn128_G1_proj.c:272-313
static void proj_add(const curve_t *C, const uint64_t *P, const uint64_t *Q, uint64_t *S) {
  int L = C->L;
  uint64_t t0[MAXL], t1[MAXL], t2[MAXL], t3[MAXL], t4[MAXL], t5[MAXL], x3[MAXL], y3[MAXL], z3[MAXL];
  fp_mul(C, PX(P), PX(Q), t0);
  fp_mul(C, PY(P), PY(Q), t1);
  fp_mul(C, PZ(P), PZ(Q), t2);
  fp_add(C, PX(P), PY(P), t3); fp_add(C, PX(Q), PY(Q), t4); fp_mul(C, t3, t4, t3);
  fp_add(C, t0, t1, t4); fp_sub(C, t3, t4, t3);                 // t3 = X1Y2 + X2Y1
  fp_add(C, PX(P), PZ(P), t4); fp_add(C, PX(Q), PZ(Q), t5); fp_mul(C, t4, t5, t4);
  fp_add(C, t0, t2, t5); fp_sub(C, t4, t5, t4);                 // t4 = X1Z2 + X2Z1
  fp_add(C, PY(P), PZ(P), t5); fp_add(C, PY(Q), PZ(Q), x3); fp_mul(C, t5, x3, t5);
  fp_add(C, t1, t2, x3); fp_sub(C, t5, x3, t5);                 // t5 = Y1Z2 + Y2Z1
  fp_scale_small(C, t2, C->b3, z3);                             // z3 = b3 t2
  fp_sub(C, t1, z3, x3); fp_add(C, t1, z3, z3);
  fp_mul(C, x3, z3, y3);
  fp_add(C, t0, t0, t1); fp_add(C, t1, t0, t1);                 // t1 = 3 t0
  fp_scale_small(C, t4, C->b3, t4);
  fp_mul(C, t1, t4, t0); fp_add(C, y3, t0, y3);
  fp_mul(C, t4, t5, t0); fp_mul(C, x3, t3, x3); fp_sub(C, x3, t0, x3);
  fp_mul(C, t1, t3, t0); fp_mul(C, z3, t5, z3); fp_add(C, z3, t0, z3);
  memcpy(PX(S), x3, 8 * L); memcpy(PY(S), y3, 8 * L); memcpy(PZ(S), z3, 8 * L);
}
// mixed addition proj + affine, EFD madd-1998-cmo with the reference's pre-checks:
// bn128_G1_proj.c:333-373
static void proj_madd(const curve_t *C, const uint64_t *P, const uint64_t *A, uint64_t *S) {
  int L = C->L;
  if (proj_is_inf(C, P)) { proj_from_affine(C, A, S); return; }
  if (aff_is_inf(C, A)) { if (S != P) memcpy(S, P, 24 * L); return; }
  uint64_t u[MAXL], uu[MAXL], v[MAXL], vv[MAXL], vvv[MAXL], Rr[MAXL], Aa[MAXL], t[MAXL];
  fp_mul(C, A + L, PZ(P), u); fp_sub(C, u, PY(P), u);
  fp_mul(C, A, PZ(P), v);     fp_sub(C, v, PX(P), v);
  if (big_is_zero(L, u) && big_is_zero(L, v)) { proj_dbl(C, P, S); return; }
  fp_sqr(C, u, uu); fp_sqr(C, v, vv); fp_mul(C, v, vv, vvv);
  fp_mul(C, vv, PX(P), Rr);
  fp_mul(C, uu, PZ(P), Aa); fp_sub(C, Aa, vvv, Aa); fp_sub(C, Aa, Rr, Aa); fp_sub(C, Aa, Rr, Aa);
  fp_sub(C, Rr, Aa, Rr);
  fp_mul(C, vvv, PY(P), t);
  fp_mul(C, PZ(P), vvv, PZ(S));
  fp_mul(C, v, Aa, PX(S));
  fp_mul(C, u, Rr, PY(S)); fp_sub(C, PY(S), t, PY(S));
}

// ------------------------------------------------------------------------------------
// Jacobian coordinates (X:Y:Z) ~ (X/Z^2, Y/Z^3), infinity = (R, R, 0): bn128_G1_jac.c:164-187
static void jac_set_inf(const curve_t *C, uint64_t *P) {
  int L = C->L; memcpy(PX(P), C->one, 8 * L); memcpy(PY(P), C->one, 8 * L); memset(PZ(P), 0, 8 * L);
}
static int jac_is_inf(const curve_t *C, const uint64_t *P) {
  int L = C->L;
  if (!(big_is_zero(L, PZ(P)) && !big_is_zero(L, PX(P)) && !big_is_zero(L, PY(P)))) return 0;
  uint64_t xx[MAXL], xxx[MAXL], yy[MAXL];
  fp_sqr(C, PX(P), xx); fp_mul(C, PX(P), xx, xxx); fp_sqr(C, PY(P), yy);
  return big_eq(L, yy, xxx);
}
static void jac_from_affine(const curve_t *C, const uint64_t *A, uint64_t *P) {
  int L = C->L;
  if (aff_is_inf(C, A)) { jac_set_inf(C, P); return; }
  memcpy(P, A, 16 * L); memcpy(PZ(P), C->one, 8 * L);
}
// bn128_G1_jac.c:120-136
static void jac_to_affine(const curve_t *C, const uint64_t *P, uint64_t *A) {
  int L = C->L;
  if (big_is_zero(L, PZ(P))) { memset(A, 0xff, 16 * L); return; }
  uint64_t zi[MAXL], zi2[MAXL], zi3[MAXL];
  fp_inv(C, PZ(P), zi); fp_sqr(C, zi, zi2); fp_mul(C, zi, zi2, zi3);
  fp_mul(C, PX(P), zi2, A);
  fp_mul(C, PY(P), zi3, A + L);
}
// EFD dbl-2007-bl (Jacobian, a = 0): bn128_G1_jac.c:236-269
static void jac_dbl(const curve_t *C, const uint64_t *P, uint64_t *Q) {
  int L = C->L;
  uint64_t XX[MAXL], YY[MAXL], YYYY[MAXL], ZZ[MAXL], S[MAXL], M[MAXL], T[MAXL], z3[MAXL], y3[MAXL];
  fp_sqr(C, PX(P), XX); fp_sqr(C, PY(P), YY); fp_sqr(C, YY, YYYY); fp_sqr(C, PZ(P), ZZ);
  fp_add(C, PX(P), YY, S); fp_sqr(C, S, S); fp_sub(C, S, XX, S); fp_sub(C, S, YYYY, S); fp_add(C, S, S, S);
  fp_add(C, XX, XX, M); fp_add(C, M, XX, M);
  fp_sqr(C, M, T); fp_sub(C, T, S, T); fp_sub(C, T, S, T);
  fp_add(C, PZ(P), PY(P), z3); fp_sqr(C, z3, z3); fp_sub(C, z3, YY, z3); fp_sub(C, z3, ZZ, z3);
  fp_sub(C, S, T, y3); fp_mul(C, y3, M, y3);
  fp_add(C, YYYY, YYYY, YYYY); fp_add(C, YYYY, YYYY, YYYY); fp_add(C, YYYY, YYYY, YYYY);
  fp_sub(C, y3, YYYY, y3);
  memcpy(PX(Q), T, 8 * L); memcpy(PY(Q), y3, 8 * L); memcpy(PZ(Q), z3, 8 * L);
}
// EFD add-2007-bl with the reference's infinity / doubling / inverse branches: bn128_G1_jac.c:278-342
static void jac_add(const curve_t *C, const uint64_t *P, const uint64_t *Q, uint64_t *S) {
  int L = C->L;
  if (jac_is_inf(C, P)) { if (S != Q) memcpy(S, Q, 24 * L); return; }
  if (jac_is_inf(C, Q)) { if (S != P) memcpy(S, P, 24 * L); return; }
  uint64_t Z1Z1[MAXL], Z2Z2[MAXL], U1[MAXL], U2[MAXL], S1[MAXL], S2[MAXL], H[MAXL], I[MAXL], J[MAXL],
      r[MAXL], V[MAXL], x3[MAXL], y3[MAXL], z3[MAXL];
  fp_sqr(C, PZ(P), Z1Z1); fp_sqr(C, PZ(Q), Z2Z2);
  fp_mul(C, PX(P), Z2Z2, U1); fp_mul(C, PX(Q), Z1Z1, U2);
  fp_mul(C, PY(P), PZ(Q), S1); fp_mul(C, S1, Z2Z2, S1);
  fp_mul(C, PY(Q), PZ(P), S2); fp_mul(C, S2, Z1Z1, S2);
  fp_sub(C, U2, U1, H);
  if (big_is_zero(L, H)) {
    if (big_eq(L, S1, S2)) jac_dbl(C, P, S); else jac_set_inf(C, S);
    return;
  }
  fp_add(C, H, H, I); fp_sqr(C, I, I);
  fp_mul(C, H, I, J);
  fp_sub(C, S2, S1, r); fp_add(C, r, r, r);
  fp_mul(C, U1, I, V);
  fp_sqr(C, r, x3); fp_sub(C, x3, J, x3); fp_sub(C, x3, V, x3); fp_sub(C, x3, V, x3);
  fp_sub(C, V, x3, y3); fp_mul(C, y3, r, y3);
  fp_mul(C, J, S1, J); fp_sub(C, y3, J, y3); fp_sub(C, y3, J, y3);
  fp_add(C, PZ(P), PZ(Q), z3); fp_sqr(C, z3, z3); fp_sub(C, z3, Z1Z1, z3); fp_sub(C, z3, Z2Z2, z3);
  fp_mul(C, z3, H, z3);
  memcpy(PX(S), x3, 8 * L); memcpy(PY(S), y3, 8 * L); memcpy(PZ(S), z3, 8 * L);
}
// EFD madd-2007-bl with the reference's branches: bn128_G1_jac.c:362-422
static void jac_madd(const curve_t *C, const uint64_t *P, const uint64_t *A, uint64_t *S) {
  int L = C->L;
  if (jac_is_inf(C, P)) { jac_from_affine(C, A, S); return; }
  if (aff_is_inf(C, A)) { if (S != P) memcpy(S, P, 24 * L); return; }
  uint64_t Z1Z1[MAXL], U2[MAXL], S2[MAXL], H[MAXL], HH[MAXL], I[MAXL], J[MAXL], r[MAXL], V[MAXL],
      x3[MAXL], y3[MAXL], z3[MAXL];
  fp_sqr(C, PZ(P), Z1Z1);
  fp_mul(C, A, Z1Z1, U2);
  fp_mul(C, A + L, PZ(P), S2); fp_mul(C, S2, Z1Z1, S2);
  fp_sub(C, U2, PX(P), H);
  fp_sub(C, S2, PY(P), r);
  if (big_is_zero(L, H)) {
    if (big_is_zero(L, r)) jac_dbl(C, P, S); else jac_set_inf(C, S);
    return;
  }
  fp_sqr(C, H, HH);
  fp_add(C, HH, HH, I); fp_add(C, I, I, I);
  fp_mul(C, H, I, J);
  fp_add(C, r, r, r);
  fp_mul(C, PX(P), I, V);
  fp_sqr(C, r, x3); fp_sub(C, x3, J, x3); fp_sub(C, x3, V, x3); fp_sub(C, x3, V, x3);
  fp_mul(C, J, PY(P), J);
  fp_sub(C, V, x3, y3); fp_mul(C, y3, r, y3); fp_sub(C, y3, J, y3); fp_sub(C, y3, J, y3);
  fp_add(C, PZ(P), H, z3); fp_sqr(C, z3, z3); fp_sub(C, z3, Z1Z1, z3); fp_sub(C, z3, HH, z3);
  memcpy(PX(S), x3, 8 * L); memcpy(PY(S), y3, 8 * L); memcpy(PZ(S), z3, 8 * L);
}

// ------------------------------------------------------------------------------------
// Pippenger MSM, unsigned c-bit windows, MSB window first.
// Template: codegen/src/Zikkurat/CodeGen/Curve/MSM.hs:86-166 ; instance bn128_G1_proj.c:506-586
typedef struct {
  void (*set_inf)(const curve_t *, uint64_t *);
  int  (*is_inf)(const curve_t *, const uint64_t *);
  void (*dbl)(const curve_t *, const uint64_t *, uint64_t *);
  void (*add)(const curve_t *, const uint64_t *, const uint64_t *, uint64_t *);
  void (*madd)(const curve_t *, const uint64_t *, const uint64_t *, uint64_t *);
  void (*to_affine)(const curve_t *, const uint64_t *, uint64_t *);
} group_t;

static const group_t PROJ = {proj_set_inf, proj_is_inf, proj_dbl, proj_add, proj_madd, proj_to_affine};
static const group_t JAC  = {jac_set_inf, jac_is_inf, jac_dbl, jac_add, jac_madd, jac_to_affine};

static void msm_variable(const curve_t *C, const group_t *G, long n, const uint64_t *expos,
                         const uint64_t *grps, uint64_t *tgt, int nlimbs, int c) {
  int L = C->L;
  int nbits = 64 * nlimbs;
  int nwindows = (nbits + c - 1) / c;
  size_t nbuckets = (size_t)1 << c;
  G->set_inf(C, tgt);
  uint64_t *S = (uint64_t *)malloc((size_t)24 * L * (nbuckets - 1));
  if (!S) abort();
#define BKT(b) (S + ((size_t)(b) - 1) * (3 * L))
  for (int K = nwindows - 1; K >= 0; K--) {
    int lo = K * c, hi = lo + c;
    if (hi > nbits) hi = nbits;
    uint64_t mask = (hi - lo >= 64) ? ~(uint64_t)0 : (((uint64_t)1 << (hi - lo)) - 1);
    for (size_t b = 1; b < nbuckets; b++) G->set_inf(C, BKT(b));
    for (long j = 0; j < n; j++) {                      // bucket accumulation :549-561
      const uint64_t *e = expos + (size_t)nlimbs * j;
      int w = lo >> 6, s = lo & 63;
      uint64_t d = e[w] >> s;
      if (s && ((hi - 1) >> 6) != w) d |= e[w + 1] << (64 - s);
      d &= mask;
      if (d) G->madd(C, BKT(d), grps + (size_t)2 * L * j, BKT(d));
    }
    uint64_t T[3 * MAXL], Rn[3 * MAXL];                 // running sums :565-574
    G->set_inf(C, T); G->set_inf(C, Rn);
    for (size_t b = nbuckets - 1; b > 0; b--) { G->add(C, T, BKT(b), T); G->add(C, Rn, T, Rn); }
    if (!G->is_inf(C, tgt)) for (int i = 0; i < c; i++) G->dbl(C, tgt, tgt);   // Horner :576-582
    G->add(C, tgt, Rn, tgt);
  }
#undef BKT
  free(S);
}

// window heuristic: bn128_G1_proj.c:596-604
static int guess_window(long n) {
  int c = (int)round(log2((double)n) - 3.5);
  if (c < 1) c = 1;
  if (c > 64) c = 64;
  return c;
}
static void msm_std(const curve_t *C, const group_t *G, long n, const uint64_t *e, const uint64_t *g,
                    uint64_t *t, int nl) {
  if (n <= 0) { G->set_inf(C, t); return; }
  msm_variable(C, G, n, e, g, t, nl, guess_window(n));
}
// bn128_G1_proj.c:629-643 (convert every scalar with Fr_mont_to_std, then the std-coefficient MSM)
static void msm_mont(const curve_t *C, const group_t *G, long n, const uint64_t *e, const uint64_t *g,
                     uint64_t *t, int nl) {
  uint64_t *s = (uint64_t *)malloc((size_t)8 * nl * (n > 0 ? n : 1));
  if (!s) abort();
  for (long i = 0; i < n; i++) fr_to_std(C, e + (size_t)nl * i, s + (size_t)nl * i);
  msm_std(C, G, n, s, g, t, nl);
  free(s);
}

// ------------------------------------------------------------------------------------
// exports
#define EXPORT __attribute__((visibility("default")))

#define DEFINE_CURVE(NAME, CV)                                                                        \
  EXPORT void zko_##NAME##_Fp_mont_mul(const uint64_t *a, const uint64_t *b, uint64_t *t) { fp_mul(&CV, a, b, t); } \
  EXPORT void zko_##NAME##_Fp_mont_add(const uint64_t *a, const uint64_t *b, uint64_t *t) { fp_add(&CV, a, b, t); } \
  EXPORT void zko_##NAME##_Fp_mont_sub(const uint64_t *a, const uint64_t *b, uint64_t *t) { fp_sub(&CV, a, b, t); } \
  EXPORT void zko_##NAME##_Fp_mont_neg(const uint64_t *a, uint64_t *t) { fp_neg(&CV, a, t); }         \
  EXPORT void zko_##NAME##_Fp_mont_inv(const uint64_t *a, uint64_t *t) { fp_inv(&CV, a, t); }         \
  EXPORT void zko_##NAME##_Fr_mont_to_std(const uint64_t *a, uint64_t *t) { fr_to_std(&CV, a, t); }   \
  EXPORT void zko_##NAME##_G1_proj_add(const uint64_t *a, const uint64_t *b, uint64_t *t) { proj_add(&CV, a, b, t); } \
  EXPORT void zko_##NAME##_G1_proj_dbl(const uint64_t *a, uint64_t *t) { proj_dbl(&CV, a, t); }       \
  EXPORT void zko_##NAME##_G1_proj_madd_proj_aff(const uint64_t *a, const uint64_t *b, uint64_t *t) { proj_madd(&CV, a, b, t); } \
  EXPORT void zko_##NAME##_G1_proj_to_affine(const uint64_t *a, uint64_t *t) { proj_to_affine(&CV, a, t); } \
  EXPORT void zko_##NAME##_G1_proj_from_affine(const uint64_t *a, uint64_t *t) { proj_from_affine(&CV, a, t); } \
  EXPORT void zko_##NAME##_G1_jac_add(const uint64_t *a, const uint64_t *b, uint64_t *t) { jac_add(&CV, a, b, t); } \
  EXPORT void zko_##NAME##_G1_jac_dbl(const uint64_t *a, uint64_t *t) { jac_dbl(&CV, a, t); }         \
  EXPORT void zko_##NAME##_G1_jac_madd_jac_aff(const uint64_t *a, const uint64_t *b, uint64_t *t) { jac_madd(&CV, a, b, t); } \
  EXPORT void zko_##NAME##_G1_jac_to_affine(const uint64_t *a, uint64_t *t) { jac_to_affine(&CV, a, t); } \
  EXPORT void zko_##NAME##_G1_proj_MSM_std_coeff_proj_out_variable(int n, const uint64_t *e, const uint64_t *g, uint64_t *t, int nl, int c) { \
    msm_variable(&CV, &PROJ, n, e, g, t, nl, c); }                                                    \
  EXPORT void zko_##NAME##_G1_proj_MSM_std_coeff_proj_out(int n, const uint64_t *e, const uint64_t *g, uint64_t *t, int nl) { \
    msm_std(&CV, &PROJ, n, e, g, t, nl); }                                                            \
  EXPORT void zko_##NAME##_G1_proj_MSM_mont_coeff_proj_out(int n, const uint64_t *e, const uint64_t *g, uint64_t *t, int nl) { \
    msm_mont(&CV, &PROJ, n, e, g, t, nl); }                                                           \
  EXPORT void zko_##NAME##_G1_proj_MSM_std_coeff_affine_out(int n, const uint64_t *e, const uint64_t *g, uint64_t *t, int nl) { \
    uint64_t tmp[3 * MAXL]; msm_std(&CV, &PROJ, n, e, g, tmp, nl); proj_to_affine(&CV, tmp, t); }     \
  EXPORT void zko_##NAME##_G1_proj_MSM_mont_coeff_affine_out(int n, const uint64_t *e, const uint64_t *g, uint64_t *t, int nl) { \
    uint64_t tmp[3 * MAXL]; msm_mont(&CV, &PROJ, n, e, g, tmp, nl); proj_to_affine(&CV, tmp, t); }    \
  EXPORT void zko_##NAME##_G1_jac_MSM_std_coeff_jac_out(int n, const uint64_t *e, const uint64_t *g, uint64_t *t, int nl) { \
    msm_std(&CV, &JAC, n, e, g, t, nl); }                                                             \
  EXPORT void zko_##NAME##_G1_jac_MSM_mont_coeff_jac_out(int n, const uint64_t *e, const uint64_t *g, uint64_t *t, int nl) { \
    msm_mont(&CV, &JAC, n, e, g, t, nl); }                                                            \
  EXPORT void zko_##NAME##_G1_jac_MSM_std_coeff_affine_out(int n, const uint64_t *e, const uint64_t *g, uint64_t *t, int nl) { \
    uint64_t tmp[3 * MAXL]; msm_std(&CV, &JAC, n, e, g, tmp, nl); jac_to_affine(&CV, tmp, t); }       \
  EXPORT void zko_##NAME##_G1_jac_MSM_mont_coeff_affine_out(int n, const uint64_t *e, const uint64_t *g, uint64_t *t, int nl) { \
    uint64_t tmp[3 * MAXL]; msm_mont(&CV, &JAC, n, e, g, tmp, nl); jac_to_affine(&CV, tmp, t); }      \
  EXPORT void zko_##NAME##_gen_chain(long n, const uint64_t *p0_aff, const uint64_t *d_aff, uint64_t *out_aff) { \
    gen_chain(&CV, n, p0_aff, d_aff, out_aff); }                                                      \
  EXPORT void zko_##NAME##_msm_threads(long n, const uint64_t *e, const uint64_t *g, uint64_t *t_aff, int mont, int nthreads) { \
    msm_threads(&CV, n, e, g, t_aff, mont, nthreads); }

// ------------------------------------------------------------------------------------
// workload helpers (not part of the reference; used by tests/bench to synthesise inputs)

// out[i] = P0 + i*D in affine Montgomery bytes.  Projective chain + one batch inversion
// (Montgomery's trick) so that 2^20..2^24 points take seconds, not minutes.
static void gen_chain(const curve_t *C, long n, const uint64_t *p0, const uint64_t *d, uint64_t *out) {
  int L = C->L;
  if (n <= 0) return;
  uint64_t *Z = (uint64_t *)malloc((size_t)8 * L * n);       // Z_i, then prefix products
  uint64_t *XY = out;                                         // X_i, Y_i stored in place
  uint64_t P[3 * MAXL];
  proj_from_affine(C, p0, P);
  for (long i = 0; i < n; i++) {
    memcpy(XY + (size_t)2 * L * i, P, 16 * L);
    memcpy(Z + (size_t)L * i, PZ(P), 8 * L);
    proj_madd(C, P, d, P);
  }
  // prefix products
  uint64_t *pre = (uint64_t *)malloc((size_t)8 * L * n);
  memcpy(pre, Z, 8 * L);
  for (long i = 1; i < n; i++) fp_mul(C, pre + (size_t)L * (i - 1), Z + (size_t)L * i, pre + (size_t)L * i);
  uint64_t inv[MAXL], zi[MAXL];
  fp_inv(C, pre + (size_t)L * (n - 1), inv);
  for (long i = n - 1; i >= 0; i--) {
    if (i > 0) { fp_mul(C, inv, pre + (size_t)L * (i - 1), zi); fp_mul(C, inv, Z + (size_t)L * i, inv); }
    else memcpy(zi, inv, 8 * L);
    fp_mul(C, XY + (size_t)2 * L * i, zi, XY + (size_t)2 * L * i);
    fp_mul(C, XY + (size_t)2 * L * i + L, zi, XY + (size_t)2 * L * i + L);
  }
  free(pre); free(Z);
}

// T contiguous shards, one pthread each, partial results combined with proj_add, then to_affine
// (the large-n oracle strategy of SURVEY.md section 8c).
typedef struct { const curve_t *C; long n; const uint64_t *e, *g; uint64_t out[3 * MAXL]; int mont; } shard_t;
static void *shard_main(void *arg) {
  shard_t *s = (shard_t *)arg;
  if (s->mont) msm_mont(s->C, &PROJ, s->n, s->e, s->g, s->out, 4);
  else         msm_std(s->C, &PROJ, s->n, s->e, s->g, s->out, 4);
  return 0;
}
static void msm_threads(const curve_t *C, long n, const uint64_t *e, const uint64_t *g, uint64_t *t_aff,
                        int mont, int nthreads) {
  int L = C->L;
  if (nthreads < 1) nthreads = 1;
  if (nthreads > 256) nthreads = 256;
  shard_t *sh = (shard_t *)calloc(nthreads, sizeof(shard_t));
  pthread_t *th = (pthread_t *)calloc(nthreads, sizeof(pthread_t));
  for (int k = 0; k < nthreads; k++) {
    long lo = n * k / nthreads, hi = n * (k + 1) / nthreads;
    sh[k].C = C; sh[k].n = hi - lo; sh[k].e = e + (size_t)4 * lo; sh[k].g = g + (size_t)2 * L * lo; sh[k].mont = mont;
    pthread_create(&th[k], 0, shard_main, &sh[k]);
  }
  uint64_t acc[3 * MAXL];
  proj_set_inf(C, acc);
  for (int k = 0; k < nthreads; k++) { pthread_join(th[k], 0); proj_add(C, acc, sh[k].out, acc); }
  proj_to_affine(C, acc, t_aff);
  free(sh); free(th);
}

DEFINE_CURVE(bn128, BN)
DEFINE_CURVE(bls12_381, BLS)

// ------------------------------------------------------------------------------------
// Array drivers for element-wise differential tests (tests/test_device_primitives.py): apply ANY function with the
// reference's element signatures -- in practice the unmodified reference's own <curve>_Fp_mont_mul & co. from
// oracle/_ref, passed in as a function pointer -- to n consecutive operands in a C loop (10^6 ctypes calls from
// Python would take longer than the GPU test budget allows).
typedef void (*zko_fn2)(const uint64_t *, uint64_t *);
typedef void (*zko_fn3)(const uint64_t *, const uint64_t *, uint64_t *);
EXPORT void zko_map2(zko_fn2 f, long n, int in_limbs, int out_limbs, const uint64_t *a, uint64_t *out) {
  for (long i = 0; i < n; i++) f(a + (size_t)i * in_limbs, out + (size_t)i * out_limbs);
}
EXPORT void zko_map3(zko_fn3 f, long n, int a_limbs, int b_limbs, int out_limbs, const uint64_t *a, const uint64_t *b, uint64_t *out) {
  for (long i = 0; i < n; i++) f(a + (size_t)i * a_limbs, b + (size_t)i * b_limbs, out + (size_t)i * out_limbs);
}
