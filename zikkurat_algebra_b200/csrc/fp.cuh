// Montgomery prime-field arithmetic on 32-bit limbs, one field element per thread, fully unrolled.
//
// Replaces (value-for-value; results are canonical < p exactly like the reference's):
//   bn128_Fp_mont_mul/_sqr        lib/cbits/curves/fields/mont/bn128_Fp_mont.c:177-199
//     = bigint256_mul             lib/cbits/bigint/bigint256.c:267-356   (schoolbook product)
//     + REDC_unsafe               lib/cbits/curves/fields/mont/bn128_Fp_mont.c:140-169
//   bn128_Fp_mont_add/_sub/_neg   lib/cbits/curves/fields/mont/bn128_Fp_mont.c:44-109
//   bn128_Fr_mont_to_std          lib/cbits/curves/fields/mont/bn128_Fr_mont.c:330-335
// and the bls12_381 twins (6x64-bit limbs there, 12x32-bit limbs here).
//
// Multiplication is the interleaved (CIOS-style) Montgomery product split into an "even" and an
// "odd" accumulator so that every 32x32->64 product lands on an aligned (lo,hi) register pair:
//   T = E + O * 2^32,  E holds limb positions 0..L-1, O holds positions 1..L.
// A row  T += a * b_i  is then two independent carry chains of L/2 products each
// (mad.lo.cc / madc.hi.cc pairs -> IMAD.WIDE.U32 with carry), and the /2^32 of each Montgomery
// step swaps the roles of the two accumulators instead of moving registers.
// Product count per multiplication: L*L (a*b) + L*L (m*p) + L (m) = 2L^2 + L  (136 | 300).
#pragma once
#include "hd.cuh"

namespace zk {

template <class P>
struct Fe {
  static constexpr int L = P::L;
  uint32_t l[P::L];
};

// acc[0..L) (+)= a[s], a[s+2], ... times b, one carry chain; CARRY_IN continues a previous chain.
// After the call the carry flag holds the carry out of acc[L-1].
template <int L, bool CARRY_IN>
ZK_HD void cmad_row(uint32_t* acc, const uint32_t* a, uint32_t b) {
#pragma unroll
  for (int j = 0; j < L; j += 2) {
    acc[j] = (j == 0 && !CARRY_IN) ? mad_lo_cc(a[j], b, acc[j]) : madc_lo_cc(a[j], b, acc[j]);
    acc[j + 1] = madc_hi_cc(a[j], b, acc[j + 1]);
  }
}

template <class P, int S>
struct ModRow {  // the constant row p[S], p[S+2], ... as an indexable object
  ZK_HD constexpr uint32_t operator[](int j) const { return P::mod(j + S); }
};

template <int L, class Row>
ZK_HD void cmad_row_const(uint32_t* acc, Row p, uint32_t m) {
#pragma unroll
  for (int j = 0; j < L; j += 2) {
    acc[j] = (j == 0) ? mad_lo_cc(p[j], m, acc[j]) : madc_lo_cc(p[j], m, acc[j]);
    acc[j + 1] = madc_hi_cc(p[j], m, acc[j + 1]);
  }
}

// r = r - p if r >= p   (r < 2p on entry)
template <class P>
ZK_HD void final_sub(uint32_t* r) {
  constexpr int L = P::L;
  uint32_t t[L];
  t[0] = sub_cc(r[0], P::mod(0));
#pragma unroll
  for (int i = 1; i < L; i++) t[i] = subc_cc(r[i], P::mod(i));
  uint32_t borrow = subc(0u, 0u);  // 0xffffffff if r < p
#pragma unroll
  for (int i = 0; i < L; i++) r[i] = borrow ? r[i] : t[i];
}

// Montgomery product, inputs canonical, output canonical.
template <class P>
ZK_HD void mont_mul_limbs(uint32_t* r, const uint32_t* a, const uint32_t* b) {
  constexpr int L = P::L;
  static_assert(L % 2 == 0, "even limb count expected");
  uint32_t A[L], B[L];  // the two accumulators; roles alternate every row
  // ---- row 0: plain products, no accumulate ----
#pragma unroll
  for (int j = 0; j < L; j += 2) {
    A[j] = mul_lo(a[j], b[0]);
    A[j + 1] = mul_hi(a[j], b[0]);
    B[j] = mul_lo(a[j + 1], b[0]);
    B[j + 1] = mul_hi(a[j + 1], b[0]);
  }
  {
    uint32_t m = mul_lo(A[0], P::INV);
    cmad_row_const<L>(B, ModRow<P, 1>(), m);  // carry out is provably 0
    cmad_row_const<L>(A, ModRow<P, 0>(), m);
    B[L - 1] = addc(B[L - 1], 0u);
  }
  // ---- rows 1..L-1 ----
#pragma unroll
  for (int i = 1; i < L; i++) {
    uint32_t* E = (i & 1) ? B : A;  // becomes the even accumulator of this row (was odd)
    uint32_t* O = (i & 1) ? A : B;  // old even accumulator; O[0] == 0, shifted down two limbs in place
    E[0] = add_cc(E[0], O[1]);
#pragma unroll
    for (int j = 0; j < L - 2; j += 2) {
      O[j] = madc_lo_cc(a[j + 1], b[i], O[j + 2]);
      O[j + 1] = madc_hi_cc(a[j + 1], b[i], O[j + 3]);
    }
    O[L - 2] = madc_lo_cc(a[L - 1], b[i], 0u);
    O[L - 1] = madc_hi(a[L - 1], b[i], 0u);
    cmad_row<L, false>(E, a, b[i]);
    O[L - 1] = addc(O[L - 1], 0u);
    uint32_t m = mul_lo(E[0], P::INV);
    cmad_row_const<L>(O, ModRow<P, 1>(), m);
    cmad_row_const<L>(E, ModRow<P, 0>(), m);
    O[L - 1] = addc(O[L - 1], 0u);
  }
  // ---- combine: result = E / 2^32 + O ----
  uint32_t* E = ((L - 1) & 1) ? B : A;
  uint32_t* O = ((L - 1) & 1) ? A : B;
  r[0] = add_cc(E[1], O[0]);
#pragma unroll
  for (int k = 1; k < L - 1; k++) r[k] = addc_cc(E[k + 1], O[k]);
  r[L - 1] = addc(O[L - 1], 0u);
  final_sub<P>(r);
}

// Two products with a common factor, r1 = a*b and r2 = a*c, row by row in turns: twice the independent carry chains in
// flight (a single product gives the scheduler two to three, and a dependent IMAD.WIDE issues ~11 cycles after its
// predecessor while the pipe could take one every 4).  Used where both products are needed anyway: the inverse peeled off a
// batch inversion (r * prefix) and the running inverse itself (r * d) in the batched-affine additions.
template <class P>
ZK_HD void mont_mul_row0(uint32_t* A, uint32_t* B, const uint32_t* a, uint32_t b0) {
  constexpr int L = P::L;
#pragma unroll
  for (int j = 0; j < L; j += 2) {
    A[j] = mul_lo(a[j], b0);
    A[j + 1] = mul_hi(a[j], b0);
    B[j] = mul_lo(a[j + 1], b0);
    B[j + 1] = mul_hi(a[j + 1], b0);
  }
  uint32_t m = mul_lo(A[0], P::INV);
  cmad_row_const<L>(B, ModRow<P, 1>(), m);  // carry out is provably 0
  cmad_row_const<L>(A, ModRow<P, 0>(), m);
  B[L - 1] = addc(B[L - 1], 0u);
}
template <class P>
ZK_HD void mont_mul_row(uint32_t* E, uint32_t* O, const uint32_t* a, uint32_t bi) {   // E: even accumulator of this row
  constexpr int L = P::L;
  E[0] = add_cc(E[0], O[1]);
#pragma unroll
  for (int j = 0; j < L - 2; j += 2) {
    O[j] = madc_lo_cc(a[j + 1], bi, O[j + 2]);
    O[j + 1] = madc_hi_cc(a[j + 1], bi, O[j + 3]);
  }
  O[L - 2] = madc_lo_cc(a[L - 1], bi, 0u);
  O[L - 1] = madc_hi(a[L - 1], bi, 0u);
  cmad_row<L, false>(E, a, bi);
  O[L - 1] = addc(O[L - 1], 0u);
  uint32_t m = mul_lo(E[0], P::INV);
  cmad_row_const<L>(O, ModRow<P, 1>(), m);
  cmad_row_const<L>(E, ModRow<P, 0>(), m);
  O[L - 1] = addc(O[L - 1], 0u);
}
template <class P>
ZK_HD void mont_mul_finish(uint32_t* r, const uint32_t* E, const uint32_t* O) {
  constexpr int L = P::L;
  r[0] = add_cc(E[1], O[0]);
#pragma unroll
  for (int k = 1; k < L - 1; k++) r[k] = addc_cc(E[k + 1], O[k]);
  r[L - 1] = addc(O[L - 1], 0u);
  final_sub<P>(r);
}
template <class P>
ZK_HD void mont_mul_pair_limbs(uint32_t* r1, uint32_t* r2, const uint32_t* a, const uint32_t* b, const uint32_t* c) {
  constexpr int L = P::L;
  uint32_t A1[L], B1[L], A2[L], B2[L];
  mont_mul_row0<P>(A1, B1, a, b[0]);
  mont_mul_row0<P>(A2, B2, a, c[0]);
#pragma unroll
  for (int i = 1; i < L; i++) {
    mont_mul_row<P>((i & 1) ? B1 : A1, (i & 1) ? A1 : B1, a, b[i]);
    mont_mul_row<P>((i & 1) ? B2 : A2, (i & 1) ? A2 : B2, a, c[i]);
  }
  mont_mul_finish<P>(r1, ((L - 1) & 1) ? B1 : A1, ((L - 1) & 1) ? A1 : B1);
  mont_mul_finish<P>(r2, ((L - 1) & 1) ? B2 : A2, ((L - 1) & 1) ? A2 : B2);
}

// ---- Karatsuba form of the product (experiment, see profiles/r2_notes.md section 16) ---------------------------------
// Every instruction that yields the high half of a 32x32-bit product occupies the IMAD pipe for 4 cycles; additions run on the
// other pipe.  One Karatsuba level computes the 2L-limb product a*b from three (L/2)^2 products instead of L^2 (108 instead of
// 144 for 12 limbs) at the price of ~140 additions, and a separate word-serial Montgomery reduction (L^2 + L products) follows.
// mul_wide_limbs: plain product of two H-limb numbers, the even/odd row scheme of mont_mul_limbs without the reduction
// (limb i of the result is final after row i).
template <int H>
ZK_HD void mul_wide_limbs(uint32_t* t, const uint32_t* a, const uint32_t* b) {
  static_assert(H % 2 == 0 && H >= 2, "even limb count expected");
  uint32_t A[H], B[H];
#pragma unroll
  for (int j = 0; j < H; j += 2) {
    A[j] = mul_lo(a[j], b[0]);
    A[j + 1] = mul_hi(a[j], b[0]);
    B[j] = mul_lo(a[j + 1], b[0]);
    B[j + 1] = mul_hi(a[j + 1], b[0]);
  }
  t[0] = A[0];
#pragma unroll
  for (int i = 1; i < H; i++) {
    uint32_t* E = (i & 1) ? B : A;
    uint32_t* O = (i & 1) ? A : B;
    E[0] = add_cc(E[0], O[1]);
#pragma unroll
    for (int j = 0; j < H - 2; j += 2) {
      O[j] = madc_lo_cc(a[j + 1], b[i], O[j + 2]);
      O[j + 1] = madc_hi_cc(a[j + 1], b[i], O[j + 3]);
    }
    O[H - 2] = madc_lo_cc(a[H - 1], b[i], 0u);
    O[H - 1] = madc_hi(a[H - 1], b[i], 0u);
    cmad_row<H, false>(E, a, b[i]);
    O[H - 1] = addc(O[H - 1], 0u);
    t[i] = E[0];
  }
  const uint32_t* E = ((H - 1) & 1) ? B : A;
  const uint32_t* O = ((H - 1) & 1) ? A : B;
  t[H] = add_cc(E[1], O[0]);
#pragma unroll
  for (int k = 1; k < H - 1; k++) t[H + k] = addc_cc(E[k + 1], O[k]);
  t[2 * H - 1] = addc(O[H - 1], 0u);
}
// d = |x - y| (H limbs), returns all-ones when x < y
template <int H>
ZK_HD uint32_t abs_diff_limbs(uint32_t* d, const uint32_t* x, const uint32_t* y) {
  d[0] = sub_cc(x[0], y[0]);
#pragma unroll
  for (int i = 1; i < H; i++) d[i] = subc_cc(x[i], y[i]);
  const uint32_t s = subc(0u, 0u);
  d[0] = sub_cc(d[0] ^ s, s);
#pragma unroll
  for (int i = 1; i < H - 1; i++) d[i] = subc_cc(d[i] ^ s, s);
  d[H - 1] = subc(d[H - 1] ^ s, s);
  return s;
}
// T (2L limbs) = a * b:  a = a0 + a1 B, b = b0 + b1 B (B = 2^(16 L)),  a0 b1 + a1 b0 = a0 b0 + a1 b1 + (a0 - a1)(b1 - b0)
template <int L>
ZK_HD void kara_mul_limbs(uint32_t* T, const uint32_t* a, const uint32_t* b) {
  constexpr int H = L / 2;
  uint32_t da[H], db[H], zm[2 * H], z1[2 * H + 1];
  const uint32_t sa = abs_diff_limbs<H>(da, a, a + H);
  const uint32_t sb = abs_diff_limbs<H>(db, b + H, b);
  mul_wide_limbs<H>(T, a, b);
  mul_wide_limbs<H>(T + 2 * H, a + H, b + H);
  mul_wide_limbs<H>(zm, da, db);
  const uint32_t neg = sa ^ sb;          // (a0 - a1)(b1 - b0) < 0
  z1[0] = add_cc(T[0], T[2 * H]);
#pragma unroll
  for (int k = 1; k < 2 * H; k++) z1[k] = addc_cc(T[k], T[2 * H + k]);
  z1[2 * H] = addc(0u, 0u);
  (void)add_cc(neg, neg);                // carry = 1 when zm is subtracted (two's complement: + ~zm + 1)
#pragma unroll
  for (int k = 0; k < 2 * H; k++) z1[k] = addc_cc(z1[k], zm[k] ^ neg);
  z1[2 * H] = addc(z1[2 * H], neg);
  T[H] = add_cc(T[H], z1[0]);
#pragma unroll
  for (int k = 1; k <= 2 * H; k++) T[H + k] = addc_cc(T[H + k], z1[k]);
#pragma unroll
  for (int k = 3 * H + 1; k < 4 * H - 1; k++) T[k] = addc_cc(T[k], 0u);
  T[4 * H - 1] = addc(T[4 * H - 1], 0u);
}
// r = T / R mod p for a 2L-limb T < p * R: word-serial reduction of the low half (rows of m * p in the even/odd scheme),
// plus the high half, one conditional subtraction.
template <class P>
ZK_HD void mont_redc_limbs(uint32_t* r, const uint32_t* T) {
  constexpr int L = P::L;
  uint32_t A[L], B[L];
#pragma unroll
  for (int k = 0; k < L; k++) A[k] = T[k];
  {
    const uint32_t m = mul_lo(A[0], P::INV);
#pragma unroll
    for (int j = 0; j < L; j += 2) {
      B[j] = mul_lo(P::mod(j + 1), m);
      B[j + 1] = mul_hi(P::mod(j + 1), m);
    }
    cmad_row_const<L>(A, ModRow<P, 0>(), m);
    B[L - 1] = addc(B[L - 1], 0u);
  }
#pragma unroll
  for (int i = 1; i < L; i++) {
    uint32_t* E = (i & 1) ? B : A;
    uint32_t* O = (i & 1) ? A : B;
    E[0] = add_cc(E[0], O[1]);
    const uint32_t m = mul_lo(E[0], P::INV);
#pragma unroll
    for (int j = 0; j < L - 2; j += 2) {
      O[j] = madc_lo_cc(P::mod(j + 1), m, O[j + 2]);
      O[j + 1] = madc_hi_cc(P::mod(j + 1), m, O[j + 3]);
    }
    O[L - 2] = madc_lo_cc(P::mod(L - 1), m, 0u);
    O[L - 1] = madc_hi(P::mod(L - 1), m, 0u);
    cmad_row_const<L>(E, ModRow<P, 0>(), m);
    O[L - 1] = addc(O[L - 1], 0u);
  }
  const uint32_t* E = ((L - 1) & 1) ? B : A;
  const uint32_t* O = ((L - 1) & 1) ? A : B;
  r[0] = add_cc(E[1], O[0]);
#pragma unroll
  for (int k = 1; k < L - 1; k++) r[k] = addc_cc(E[k + 1], O[k]);
  r[L - 1] = addc(O[L - 1], 0u);
  r[0] = add_cc(r[0], T[L]);
#pragma unroll
  for (int k = 1; k < L - 1; k++) r[k] = addc_cc(r[k], T[L + k]);
  r[L - 1] = addc(r[L - 1], T[2 * L - 1]);
  final_sub<P>(r);
}
template <class P>
ZK_HD void mont_mul_kara_limbs(uint32_t* r, const uint32_t* a, const uint32_t* b) {
  uint32_t T[2 * P::L];
  kara_mul_limbs<P::L>(T, a, b);
  mont_redc_limbs<P>(r, T);
}

// Fused r = (a*b + c*d) * R^-1 mod p with ONE interleaved reduction: 3L^2 + L products instead of the
// 4L^2 + 2L of two separate multiplications (used for Y3 = R*(Q-X3) + (-Y1)*PPP in every group addition).
// Inputs canonical.  Row bound: T_i < 3p(1 + 2^-32), and T + 3*2^32*p < 2^(32(L+1)) needs 3p < 2^(32L)
// (0.57 for BN254, 0.31 for BLS12-381); final value (ab + cd + Mp)/R < p(1 + 2p/R) < 2p: one subtraction.
template <class P>
ZK_HD void mont_mul2_limbs(uint32_t* r, const uint32_t* a, const uint32_t* b, const uint32_t* c, const uint32_t* d) {
  static_assert(P::THREE_MOD_FITS, "fused product needs 3*mod < 2^(32L) (true for both Fp, false for BLS12-381 Fr)");
  constexpr int L = P::L;
  uint32_t A[L], B[L];
#pragma unroll
  for (int j = 0; j < L; j += 2) {
    A[j] = mul_lo(a[j], b[0]);
    A[j + 1] = mul_hi(a[j], b[0]);
    B[j] = mul_lo(a[j + 1], b[0]);
    B[j + 1] = mul_hi(a[j + 1], b[0]);
  }
  cmad_row<L, false>(B, c + 1, d[0]);  // odd limbs of c; carry out is provably 0
  cmad_row<L, false>(A, c, d[0]);
  B[L - 1] = addc(B[L - 1], 0u);
  {
    uint32_t m = mul_lo(A[0], P::INV);
    cmad_row_const<L>(B, ModRow<P, 1>(), m);
    cmad_row_const<L>(A, ModRow<P, 0>(), m);
    B[L - 1] = addc(B[L - 1], 0u);
  }
#pragma unroll
  for (int i = 1; i < L; i++) {
    uint32_t* E = (i & 1) ? B : A;
    uint32_t* O = (i & 1) ? A : B;
    E[0] = add_cc(E[0], O[1]);
#pragma unroll
    for (int j = 0; j < L - 2; j += 2) {
      O[j] = madc_lo_cc(a[j + 1], b[i], O[j + 2]);
      O[j + 1] = madc_hi_cc(a[j + 1], b[i], O[j + 3]);
    }
    O[L - 2] = madc_lo_cc(a[L - 1], b[i], 0u);
    O[L - 1] = madc_hi(a[L - 1], b[i], 0u);
    cmad_row<L, false>(O, c + 1, d[i]);
    cmad_row<L, false>(E, a, b[i]);
    O[L - 1] = addc(O[L - 1], 0u);
    cmad_row<L, false>(E, c, d[i]);
    O[L - 1] = addc(O[L - 1], 0u);
    uint32_t m = mul_lo(E[0], P::INV);
    cmad_row_const<L>(O, ModRow<P, 1>(), m);
    cmad_row_const<L>(E, ModRow<P, 0>(), m);
    O[L - 1] = addc(O[L - 1], 0u);
  }
  uint32_t* E = ((L - 1) & 1) ? B : A;
  uint32_t* O = ((L - 1) & 1) ? A : B;
  r[0] = add_cc(E[1], O[0]);
#pragma unroll
  for (int k = 1; k < L - 1; k++) r[k] = addc_cc(E[k + 1], O[k]);
  r[L - 1] = addc(O[L - 1], 0u);
  final_sub<P>(r);
}

template <class P>
ZK_HD Fe<P> fe_mul(const Fe<P>& a, const Fe<P>& b) {
  Fe<P> r;
  mont_mul_limbs<P>(r.l, a.l, b.l);
  return r;
}
// Montgomery squaring.  Same row structure as mont_mul_limbs, but row i only adds
//   a_i * ( a_i*2^(32i) + 2*sum_{j>i} a_j*2^(32j) )
// i.e. L-i products instead of L (the cross terms a_i*a_j are taken once, doubled): L(L+1)/2 + L^2 + L
// products (108 | 234) instead of 2L^2 + L (136 | 300).  The doubled multiplicand 2a is precomputed as
// 32-bit limbs d_j = (a_j<<1)|(a_{j-1}>>31); the limb right above the diagonal must not contain the bit
// carried up from a_i, hence d_{i+1} & ~1.  a < p < 2^(32L-2), so 2a has no limb L.
// Where a row has no product the chain still has to carry (and, for the accumulator that is shifted down
// two limbs per row, move) the limb: addc.cc with 0 takes the place of madc.  Row bound as for
// mont_mul2_limbs: each row adds a_i*W + m*p with W <= 2a < 2p, T stays below 2^(32(L+1)) because 3p < 2^(32L).
template <class P>
ZK_HD void mont_sqr_limbs(uint32_t* r, const uint32_t* a) {
  static_assert(P::THREE_MOD_FITS, "dedicated squaring needs 3*mod < 2^(32L)");
  constexpr int L = P::L;
  uint32_t A[L], B[L], d[L];
  d[0] = a[0] << 1;  // unused as a multiplicand (row 0 uses a_0 itself), kept for uniform indexing
#pragma unroll
  for (int j = 1; j < L; j++) d[j] = (a[j] << 1) | (a[j - 1] >> 31);
  // multiplicand limb j of row i (j >= i):  a_i on the diagonal, d_{i+1} without its carried-in bit, d_j above
#define ZK_SQR_W(i, j) ((j) == (i) ? a[(j)] : ((j) == (i) + 1 ? (d[(j)] & 0xfffffffeu) : d[(j)]))
  // ---- row 0: full row ----
#pragma unroll
  for (int j = 0; j < L; j += 2) {
    A[j] = mul_lo(ZK_SQR_W(0, j), a[0]);
    A[j + 1] = mul_hi(ZK_SQR_W(0, j), a[0]);
    B[j] = mul_lo(ZK_SQR_W(0, j + 1), a[0]);
    B[j + 1] = mul_hi(ZK_SQR_W(0, j + 1), a[0]);
  }
  {
    uint32_t m = mul_lo(A[0], P::INV);
    cmad_row_const<L>(B, ModRow<P, 1>(), m);
    cmad_row_const<L>(A, ModRow<P, 0>(), m);
    B[L - 1] = addc(B[L - 1], 0u);
  }
#pragma unroll
  for (int i = 1; i < L; i++) {
    uint32_t* E = (i & 1) ? B : A;
    uint32_t* O = (i & 1) ? A : B;
    E[0] = add_cc(E[0], O[1]);
    // odd positions t = j + 1 (pairs (t, t+1) live in O[j], O[j+1] after the two-limb shift)
#pragma unroll
    for (int j = 0; j < L - 2; j += 2) {
      if (j + 1 >= i) {
        O[j] = madc_lo_cc(ZK_SQR_W(i, j + 1), a[i], O[j + 2]);
        O[j + 1] = madc_hi_cc(ZK_SQR_W(i, j + 1), a[i], O[j + 3]);
      } else {
        O[j] = addc_cc(O[j + 2], 0u);
        O[j + 1] = addc_cc(O[j + 3], 0u);
      }
    }
    O[L - 2] = madc_lo_cc(ZK_SQR_W(i, L - 1), a[i], 0u);   // position L-1 >= i always
    O[L - 1] = madc_hi(ZK_SQR_W(i, L - 1), a[i], 0u);
    // even positions t = j >= i
    bool started = false;
#pragma unroll
    for (int j = 0; j < L; j += 2) {
      if (j >= i) {
        E[j] = started ? madc_lo_cc(ZK_SQR_W(i, j), a[i], E[j]) : mad_lo_cc(ZK_SQR_W(i, j), a[i], E[j]);
        E[j + 1] = madc_hi_cc(ZK_SQR_W(i, j), a[i], E[j + 1]);
        started = true;
      }
    }
    if (started) O[L - 1] = addc(O[L - 1], 0u);
    uint32_t m = mul_lo(E[0], P::INV);
    cmad_row_const<L>(O, ModRow<P, 1>(), m);
    cmad_row_const<L>(E, ModRow<P, 0>(), m);
    O[L - 1] = addc(O[L - 1], 0u);
  }
#undef ZK_SQR_W
  uint32_t* E = ((L - 1) & 1) ? B : A;
  uint32_t* O = ((L - 1) & 1) ? A : B;
  r[0] = add_cc(E[1], O[0]);
#pragma unroll
  for (int k = 1; k < L - 1; k++) r[k] = addc_cc(E[k + 1], O[k]);
  r[L - 1] = addc(O[L - 1], 0u);
  final_sub<P>(r);
}

template <class P>
ZK_HD Fe<P> fe_sqr(const Fe<P>& a) {
  Fe<P> r;
  if constexpr (P::THREE_MOD_FITS) mont_sqr_limbs<P>(r.l, a.l);
  else mont_mul_limbs<P>(r.l, a.l, a.l);   // BLS12-381 Fr (r ~ 0.45 * 2^256): the row bound does not hold
  return r;
}

// Out-of-line multiplication: ONE copy of the unrolled product per kernel instead of one per call site.
// Operands travel in registers (no stack traffic); used where the fully inlined group operation would
// not fit the instruction cache (k_accumulate: 10 multiplications per insertion).
#if defined(__CUDACC__)
// Build-time experiment switch (tools/build_variant.py ... -DZK_KARA=1): the out-of-line general products of 12-limb fields
// through mont_mul_kara_limbs.
#ifndef ZK_KARA
#define ZK_KARA 0
#endif
template <class P>
__device__ __noinline__ Fe<P> fe_mul_call(Fe<P> a, Fe<P> b) {
  Fe<P> r;
  if constexpr (ZK_KARA && P::L >= 12) mont_mul_kara_limbs<P>(r.l, a.l, b.l);
  else mont_mul_limbs<P>(r.l, a.l, b.l);
  return r;
}
template <class P>
__device__ __noinline__ Fe<P> fe_sqr_call(Fe<P> a) {
  return fe_sqr<P>(a);
}
template <class P>
struct FePair { Fe<P> u, v; };
template <class P>
__device__ __noinline__ FePair<P> fe_mul_pair_call(Fe<P> a, Fe<P> b, Fe<P> c) {
  FePair<P> r;
  if constexpr (ZK_KARA && P::L >= 12) {
    mont_mul_kara_limbs<P>(r.u.l, a.l, b.l);
    mont_mul_kara_limbs<P>(r.v.l, a.l, c.l);
  } else {
    mont_mul_pair_limbs<P>(r.u.l, r.v.l, a.l, b.l, c.l);
  }
  return r;
}
template <class P>
__device__ __noinline__ Fe<P> fe_mul2_call(Fe<P> a, Fe<P> b, Fe<P> c, Fe<P> d) {
  Fe<P> r;
  mont_mul2_limbs<P>(r.l, a.l, b.l, c.l, d.l);
  return r;
}
#else
template <class P>
inline Fe<P> fe_mul_call(Fe<P> a, Fe<P> b) { return fe_mul<P>(a, b); }
template <class P>
inline Fe<P> fe_sqr_call(Fe<P> a) { return fe_sqr<P>(a); }
template <class P>
struct FePair { Fe<P> u, v; };
template <class P>
inline FePair<P> fe_mul_pair_call(Fe<P> a, Fe<P> b, Fe<P> c) {
  FePair<P> r;
  mont_mul_pair_limbs<P>(r.u.l, r.v.l, a.l, b.l, c.l);
  return r;
}
template <class P>
inline Fe<P> fe_mul2_call(Fe<P> a, Fe<P> b, Fe<P> c, Fe<P> d) {
  Fe<P> r;
  mont_mul2_limbs<P>(r.l, a.l, b.l, c.l, d.l);
  return r;
}
#endif

template <class P>
ZK_HD Fe<P> fe_add(const Fe<P>& a, const Fe<P>& b) {
  constexpr int L = P::L;
  Fe<P> r;
  r.l[0] = add_cc(a.l[0], b.l[0]);
#pragma unroll
  for (int i = 1; i < L - 1; i++) r.l[i] = addc_cc(a.l[i], b.l[i]);
  r.l[L - 1] = addc(a.l[L - 1], b.l[L - 1]);  // p < 2^(32L-1): no carry out
  final_sub<P>(r.l);
  return r;
}

template <class P>
ZK_HD Fe<P> fe_sub(const Fe<P>& a, const Fe<P>& b) {
  constexpr int L = P::L;
  Fe<P> r;
  r.l[0] = sub_cc(a.l[0], b.l[0]);
#pragma unroll
  for (int i = 1; i < L; i++) r.l[i] = subc_cc(a.l[i], b.l[i]);
  uint32_t borrow = subc(0u, 0u);  // all-ones when a < b
  r.l[0] = add_cc(r.l[0], P::mod(0) & borrow);
#pragma unroll
  for (int i = 1; i < L - 1; i++) r.l[i] = addc_cc(r.l[i], P::mod(i) & borrow);
  r.l[L - 1] = addc(r.l[L - 1], P::mod(L - 1) & borrow);
  return r;
}

// a*b + c*d (see mont_mul2_limbs)
template <class P>
ZK_HD Fe<P> fe_mul2(const Fe<P>& a, const Fe<P>& b, const Fe<P>& c, const Fe<P>& d) {
  Fe<P> r;
  mont_mul2_limbs<P>(r.l, a.l, b.l, c.l, d.l);
  return r;
}

template <class P>
ZK_HD bool fe_is_zero(const Fe<P>& a) {
  uint32_t o = a.l[0];
#pragma unroll
  for (int i = 1; i < P::L; i++) o |= a.l[i];
  return o == 0;
}

template <class P>
ZK_HD Fe<P> fe_neg(const Fe<P>& a) {  // 0 -> 0, else p - a
  constexpr int L = P::L;
  Fe<P> r;
  uint32_t nz = fe_is_zero<P>(a) ? 0u : 0xffffffffu;
  r.l[0] = sub_cc(P::mod(0) & nz, a.l[0]);
#pragma unroll
  for (int i = 1; i < L - 1; i++) r.l[i] = subc_cc(P::mod(i) & nz, a.l[i]);
  r.l[L - 1] = subc(P::mod(L - 1) & nz, a.l[L - 1]);
  return r;
}

template <class P>
ZK_HD Fe<P> fe_dbl(const Fe<P>& a) {
  return fe_add<P>(a, a);
}

template <class P>
ZK_HD Fe<P> fe_zero() {
  Fe<P> r;
#pragma unroll
  for (int i = 0; i < P::L; i++) r.l[i] = 0;
  return r;
}
template <class P>
ZK_HD Fe<P> fe_one() {
  Fe<P> r;
#pragma unroll
  for (int i = 0; i < P::L; i++) r.l[i] = P::one(i);
  return r;
}
template <class P>
ZK_HD bool fe_eq(const Fe<P>& a, const Fe<P>& b) {
  uint32_t o = 0;
#pragma unroll
  for (int i = 0; i < P::L; i++) o |= a.l[i] ^ b.l[i];
  return o == 0;
}

// Montgomery -> standard representation: one REDC of (a, 0), i.e. a * 1 * R^-1.
template <class P>
ZK_HD Fe<P> fe_from_mont(const Fe<P>& a) {
  Fe<P> one_std = fe_zero<P>();
  one_std.l[0] = 1;
  return fe_mul<P>(a, one_std);
}

// a^(p-2) by square-and-multiply over the constant exponent (kept as a cross-check of fe_inv).
template <class P>
ZK_HD Fe<P> fe_inv_fermat(const Fe<P>& a) {
  constexpr int L = P::L;
  Fe<P> acc = fe_one<P>();
  bool started = false;
  for (int i = 32 * L - 1; i >= 0; i--) {
    // exponent p - 2: p is odd and p mod 2^32 >= 2 for both primes, so only limb 0 changes
    uint32_t w = P::mod(i >> 5) - ((i >> 5) == 0 ? 2u : 0u);
    uint32_t bit = (w >> (i & 31)) & 1u;
    if (started) acc = fe_sqr<P>(acc);
    if (bit) {
      acc = started ? fe_mul<P>(acc, a) : a;
      started = true;
    }
  }
  return acc;
}

// ---- inversion by the binary extended Euclidean algorithm -------------------------------------------
// Same method as the reference (bn128_Fp_std_inv, lib/cbits/curves/fields/std/bn128_Fp_std.c:252-315,
// followed by a multiplication by R^3, lib/cbits/curves/fields/mont/bn128_Fp_mont.c:201-204): the
// Montgomery residue x = a*R is inverted as a plain integer mod p, and x^-1 * R^3 * R^-1 = a^-1 * R.
// ~2*bits shift/subtract steps on L-limb integers: about 6x shorter than the Fermat chain for one thread.
template <int L>
ZK_HD void big_shr1(uint32_t* a, uint32_t top_in) {  // a = (top_in:a) >> 1
#pragma unroll
  for (int i = 0; i < L - 1; i++) a[i] = (a[i] >> 1) | (a[i + 1] << 31);
  a[L - 1] = (a[L - 1] >> 1) | (top_in << 31);
}
template <class P>
ZK_HD void halve_mod(uint32_t* x) {  // x = x/2 mod p  (x < p)
  constexpr int L = P::L;
  uint32_t odd = (x[0] & 1u) ? 0xffffffffu : 0u;
  x[0] = add_cc(x[0], P::mod(0) & odd);
#pragma unroll
  for (int i = 1; i < L; i++) x[i] = addc_cc(x[i], P::mod(i) & odd);
  uint32_t top = addc(0u, 0u);
  big_shr1<L>(x, top);
}
template <int L>
ZK_HD bool big_is_one(const uint32_t* a) {
  uint32_t o = a[0] ^ 1u;
#pragma unroll
  for (int i = 1; i < L; i++) o |= a[i];
  return o == 0;
}
// a -= b, returns true when it borrowed
template <int L>
ZK_HD bool big_sub_borrow(uint32_t* a, const uint32_t* b) {
  a[0] = sub_cc(a[0], b[0]);
#pragma unroll
  for (int i = 1; i < L; i++) a[i] = subc_cc(a[i], b[i]);
  return subc(0u, 0u) != 0u;
}
template <int L>
ZK_HD bool big_lt(const uint32_t* a, const uint32_t* b) {  // a < b
  uint32_t t = sub_cc(a[0], b[0]);
#pragma unroll
  for (int i = 1; i < L; i++) t = subc_cc(a[i], b[i]);
  (void)t;
  return subc(0u, 0u) != 0u;
}
template <class P>
ZK_HD Fe<P> fe_inv_euclid(const Fe<P>& a) {  // a != 0, canonical; returns the Montgomery form of 1/a
  constexpr int L = P::L;
  uint32_t u[L], v[L];
  Fe<P> x1 = fe_zero<P>(), x2 = fe_zero<P>();
  x1.l[0] = 1;
#pragma unroll
  for (int i = 0; i < L; i++) { u[i] = a.l[i]; v[i] = P::mod(i); }
  for (int guard = 0; guard < 4 * 32 * L; guard++) {
    if (big_is_one<L>(u) || big_is_one<L>(v)) break;
    while ((u[0] & 1u) == 0) { big_shr1<L>(u, 0u); halve_mod<P>(x1.l); }
    while ((v[0] & 1u) == 0) { big_shr1<L>(v, 0u); halve_mod<P>(x2.l); }
    if (!big_lt<L>(u, v)) { big_sub_borrow<L>(u, v); x1 = fe_sub<P>(x1, x2); }
    else { big_sub_borrow<L>(v, u); x2 = fe_sub<P>(x2, x1); }
  }
  Fe<P> r = big_is_one<L>(u) ? x1 : x2;
  Fe<P> r3;
#pragma unroll
  for (int i = 0; i < L; i++) r3.l[i] = P::r3(i);
  return fe_mul<P>(r, r3);
}


// ---- inversion, fast path: Kaliski's almost-Montgomery inverse ------------------------------------------------
// Phase 1 is the same binary gcd walk, but the cofactors r, s are only shifted and added -- no modular halving
// inside the loop -- so an iteration is one subtraction and one addition (independent carry chains) instead of
// six dependent ones.  It ends with  x = a^-1 * 2^k mod p,  bits <= k <= 2*bits.  Phase 2 removes 2^k and
// restores the Montgomery scale with three multiplications:
//   x * R^3/R = x*R^2;   then  * 2^e1 / R  and  * 2^e2 / R  with  e1 + e2 = 2*32L - k   =>   a^-1 * R^2 (a = residue).
// The result is the canonical inverse, bit-identical to fe_inv_euclid (tests/test_host_emul.py).
template <int L>
ZK_HD void big_shl1(uint32_t* a) {
#pragma unroll
  for (int i = L - 1; i > 0; i--) a[i] = (a[i] << 1) | (a[i - 1] >> 31);
  a[0] <<= 1;
}
template <int L>
ZK_HD bool big_is_zero(const uint32_t* a) {
  uint32_t o = a[0];
#pragma unroll
  for (int i = 1; i < L; i++) o |= a[i];
  return o == 0;
}
template <class P>
ZK_HD Fe<P> fe_pow2(int e) {  // 2^e as a plain integer, e < 32L
  Fe<P> r;
#pragma unroll
  for (int i = 0; i < P::L; i++) r.l[i] = (i == (e >> 5)) ? (1u << (e & 31)) : 0u;
  return r;
}
// multi-bit shifts, 1 <= t <= 31
ZK_HD uint32_t funnel_r(uint32_t lo, uint32_t hi, int t) {
#if defined(__CUDA_ARCH__)
  return __funnelshift_r(lo, hi, t);
#else
  return (lo >> t) | (hi << (32 - t));
#endif
}
ZK_HD int ctz32(uint32_t x) {  // x != 0
#if defined(__CUDA_ARCH__)
  return __ffs((int)x) - 1;
#else
  return __builtin_ctz(x);
#endif
}
template <int L>
ZK_HD void big_shr(uint32_t* a, int t) {
#pragma unroll
  for (int i = 0; i < L - 1; i++) a[i] = funnel_r(a[i], a[i + 1], t);
  a[L - 1] >>= t;
}
template <int L>
ZK_HD void big_shl(uint32_t* a, int t) {
#pragma unroll
  for (int i = L - 1; i > 0; i--) a[i] = funnel_r(a[i - 1], a[i], 32 - t);
  a[0] <<= t;
}
// trailing zero bits of a non-zero multi-limb value, capped at 31 (the caller loops)
ZK_HD int low_zeros(uint32_t limb0) { return limb0 ? ctz32(limb0) : 31; }

template <class P>
ZK_HD Fe<P> fe_inv(const Fe<P>& a) {  // a != 0, canonical; returns the Montgomery form of 1/a
  constexpr int L = P::L;
  uint32_t u[L], v[L], r[L], s[L];
#pragma unroll
  for (int i = 0; i < L; i++) { u[i] = P::mod(i); v[i] = a.l[i]; r[i] = 0; s[i] = 0; }
  s[0] = 1;
  int k = 0;
  // Every trip ends with u and v odd again: all the halvings that follow a subtraction are done at once
  // (count trailing zeros), so a trip is one subtraction, one addition and two multi-bit shifts.
  for (int guard = 0; guard < 2 * P::BITS + 2; guard++) {
    if (big_is_zero<L>(v)) break;
    if ((u[0] & 1u) == 0) {
      const int t = low_zeros(u[0]);
      big_shr<L>(u, t); big_shl<L>(s, t); k += t;
      continue;
    }
    if ((v[0] & 1u) == 0) {
      const int t = low_zeros(v[0]);
      big_shr<L>(v, t); big_shl<L>(r, t); k += t;
      continue;
    }
    uint32_t d[L], rs[L];   // u - v and r + s: independent carry chains
    d[0] = sub_cc(u[0], v[0]);
#pragma unroll
    for (int i = 1; i < L; i++) d[i] = subc_cc(u[i], v[i]);
    const bool u_lt_v = subc(0u, 0u) != 0u;
    rs[0] = add_cc(r[0], s[0]);
#pragma unroll
    for (int i = 1; i < L - 1; i++) rs[i] = addc_cc(r[i], s[i]);
    rs[L - 1] = addc(r[L - 1], s[L - 1]);
    if (!u_lt_v && !big_is_zero<L>(d)) {   // u > v:  u = (u - v) / 2^t,  r = r + s,  s = s * 2^t
      const int t = low_zeros(d[0]);
#pragma unroll
      for (int i = 0; i < L; i++) { u[i] = d[i]; r[i] = rs[i]; }
      big_shr<L>(u, t); big_shl<L>(s, t); k += t;
    } else {                               // v >= u: v = (v - u) / 2^t,  s = s + r,  r = r * 2^t
      v[0] = sub_cc(v[0], u[0]);
#pragma unroll
      for (int i = 1; i < L; i++) v[i] = subc_cc(v[i], u[i]);
#pragma unroll
      for (int i = 0; i < L; i++) s[i] = rs[i];
      if (big_is_zero<L>(v)) { big_shl1<L>(r); k += 1; break; }   // u == v == 1: the last step halves a zero
      const int t = low_zeros(v[0]);
      big_shr<L>(v, t); big_shl<L>(r, t); k += t;
    }
  }
  // r < 2p:  x = p - (r mod p)
  uint32_t pm[L];
#pragma unroll
  for (int i = 0; i < L; i++) pm[i] = P::mod(i);
  if (!big_lt<L>(r, pm)) big_sub_borrow<L>(r, pm);
  Fe<P> x;
  x.l[0] = sub_cc(pm[0], r[0]);
#pragma unroll
  for (int i = 1; i < L - 1; i++) x.l[i] = subc_cc(pm[i], r[i]);
  x.l[L - 1] = subc(pm[L - 1], r[L - 1]);
  Fe<P> r3;
#pragma unroll
  for (int i = 0; i < L; i++) r3.l[i] = P::r3(i);
  const int e_total = 2 * 32 * L - k;
  const int e1 = e_total < P::BITS - 1 ? e_total : P::BITS - 1;
  const int e2 = e_total - e1;
  Fe<P> t = fe_mul<P>(x, r3);
  t = fe_mul<P>(t, fe_pow2<P>(e1));
  return fe_mul<P>(t, fe_pow2<P>(e2));
}

}  // namespace zk
