// FFT of group elements (G1 in projective or Jacobian input form, G2 projective) -- scope row 8f.4 (KZG setup: examples/KZG.hs:55, lib/src/.../G1/Affine.hs:152-157).
// Same semantics as the reference's recursive routines
//   <curve>_G1_proj_fft_forward / _inverse      lib/cbits/curves/g1/proj/bn128_G1_proj.c:678-789
//   forward:  tgt[k] = sum_j gen^(j*k) * src[j]            inverse:  tgt[j] = N^-1 * sum_k gen^(-j*k) * src[k]
// on N = 2^m projective points, natural order in and out, results NORMALISED ((x, y, 1) or the infinity (0, 1, 0)
// exactly like bn128_G1_proj_normalize, :75-96; the Jacobian and G2 twins normalise the same way:
// bn128_G1_jac.c:62-85, bn128_G2_proj.c:70-89), hence bit-comparable.
//
// Radix-2 decimation in time on XYZZ points held in global memory: one launch per stage, one 4-lane TEAM per butterfly
// (a, b) -> (a + w*b, a - w*b); w*b is a 4-bit fixed-window scalar multiplication (the reference does the same per
// butterfly with bn128_G1_proj_scl_Fr_mont).  Twiddles come from the Fr table of ntt.cu (w^-i = -w^(N/2-i)).
// The work is N/2 * log2 N scalar multiplications (~3 300 Fp multiplications each): IMAD-bound, no data reuse to stage.
// A stage has only N/2 independent scalar multiplications (8192 at 2^14: a fifth of the GPU's resident threads), each a
// chain of ~330 dependent group operations, so a butterfly is worked on by the 4 lanes of a team (ec_team.cuh: a
// doubling is 3 multiplication rounds instead of 9 dependent multiplications, an addition 4 instead of 14).
#include <cuda_runtime.h>

#include "ec_team.cuh"
#include "gfft.cuh"
#include "glv.cuh"
#include "msm_common.cuh"
#include "ntt.cuh"

namespace zk {

// one out-of-line copy of each group operation for this translation unit (team versions: all 4 lanes of a team call
// them with identical operands)
template <class P>
__device__ __noinline__ void gf_add(const Team& tm, Xyzz<P>& a, const Xyzz<P>& b) { a = xyzz_add_team<P>(tm, a, b); }
template <class P>
__device__ __noinline__ void gf_dbl(const Team& tm, Xyzz<P>& a) { a = xyzz_dbl_team<P>(tm, a); }

// k * p for a standard-form 256-bit scalar k (8 x u32), 4-bit fixed windows, table of 1p..15p in local memory
template <class P>
__device__ __noinline__ Xyzz<P> xyzz_scalar_mul(const Team& tm, const Xyzz<P>& p, const uint32_t* k) {
  Xyzz<P> tab[15];
  tab[0] = p;
  for (int i = 1; i < 15; i++) {           // tab[i] = (i+1) * p
    if (i & 1) { tab[i] = tab[i >> 1]; gf_dbl<P>(tm, tab[i]); }
    else { tab[i] = tab[i - 1]; gf_add<P>(tm, tab[i], p); }
  }
  Xyzz<P> acc = xyzz_inf<P>();
  for (int w = 63; w >= 0; w--) {
    for (int d = 0; d < 4; d++) gf_dbl<P>(tm, acc);   // no-op while acc is infinity
    uint32_t dig = (k[w >> 3] >> ((w & 7) * 4)) & 15u;
    if (dig) gf_add<P>(tm, acc, tab[dig - 1]);
  }
  return acc;
}

// The same product through the curve endomorphism (G1 only, glv.cuh): k = k1 + k2 lambda with |k1|, |k2| < 2^127, so
// k p = k1 p + k2 phi(p) by ONE joint chain of 128 doublings over 2-bit windows of both halves (table i p1 + j p2,
// i, j < 4) instead of 256 doublings.  Same group element; precondition and switch as for the MSM (zkb200_set_glv).
template <class C>
__device__ __noinline__ Xyzz<typename C::Fp> xyzz_scalar_mul_glv(const Team& tm, const Xyzz<typename C::Fp>& p, const uint32_t* k) {
  using P = typename C::Fp;
  using G = typename GlvOf<C>::type;
  uint32_t k1[4], k2[4];
  bool n1, n2;
  glv_decompose<G>(k, k1, n1, k2, n2);
  Xyzz<P> tab[16];
  tab[0] = xyzz_inf<P>();
  tab[1] = p;
  tab[4] = p;
  tab[4].X = glv_beta_x<P, G>(p.X);                // phi(x, y) = (beta x, y); x = X / ZZ
  if (n1) tab[1].Y = fe_neg<P>(tab[1].Y);
  if (n2) tab[4].Y = fe_neg<P>(tab[4].Y);
  tab[2] = tab[1]; gf_dbl<P>(tm, tab[2]);
  tab[3] = tab[2]; gf_add<P>(tm, tab[3], tab[1]);
  tab[8] = tab[4]; gf_dbl<P>(tm, tab[8]);
  tab[12] = tab[8]; gf_add<P>(tm, tab[12], tab[4]);
  for (int j = 1; j < 4; j++)
    for (int i = 1; i < 4; i++) { tab[4 * j + i] = tab[4 * j]; gf_add<P>(tm, tab[4 * j + i], tab[i]); }
  Xyzz<P> acc = xyzz_inf<P>();
  for (int w = 63; w >= 0; w--) {
    gf_dbl<P>(tm, acc);                             // no-ops while acc is infinity
    gf_dbl<P>(tm, acc);
    const int sh = (w & 15) * 2;
    const uint32_t idx = (((k2[w >> 4] >> sh) & 3u) << 2) | ((k1[w >> 4] >> sh) & 3u);
    if (idx) gf_add<P>(tm, acc, tab[idx]);
  }
  return acc;
}
template <class C>
__device__ __forceinline__ Xyzz<typename C::Fp> gfft_scalar_mul(const Team& tm, const Xyzz<typename C::Fp>& p, const uint32_t* k, int glv) {
  if constexpr (GlvOf<C>::available) {
    if (glv) return xyzz_scalar_mul_glv<C>(tm, p, k);
  }
  return xyzz_scalar_mul<typename C::Fp>(tm, p, k);
}

__device__ __forceinline__ size_t gfft_bitrev(size_t x, int bits) {
  return bits == 0 ? 0 : (size_t)(__brevll((unsigned long long)x) >> (64 - bits));
}

// projective (X:Y:Z) records -> XYZZ, written to the bit-reversed position
template <class C>
__global__ void __launch_bounds__(128) k_gfft_load(const uint32_t* __restrict__ src, int m, int jac, XyzzMem<typename C::Fp>* __restrict__ dst) {
  using P = typename C::Fp;
  constexpr int L = P::L;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ((size_t)1 << m)) return;
  const uint32_t* s = src + i * 3 * L;
  Fe<P> X, Y, Z;
  for (int k = 0; k < L; k++) { X.l[k] = s[k]; Y.l[k] = s[L + k]; Z.l[k] = s[2 * L + k]; }
  store_xyzz<P>(dst + gfft_bitrev(i, m), jac ? xyzz_from_jac<P>(X, Y, Z) : xyzz_from_proj<P>(X, Y, Z));
}

// stage s (1-based): butterflies at distance 2^(s-1) inside blocks of 2^s
template <class C>
__global__ void __launch_bounds__(128)
k_gfft_stage(XyzzMem<typename C::Fp>* __restrict__ data, const uint32_t* __restrict__ table, int m, int s, int inverse, int glv) {
  using P = typename C::Fp;
  using F = typename C::Fr;
  size_t t = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 2;   // butterfly of this team
  const size_t half = (size_t)1 << (m - 1);
  if (t >= half) return;
  const Team tm;
  const size_t j = t & (((size_t)1 << (s - 1)) - 1);
  const size_t ia = ((t >> (s - 1)) << s) + j, ib = ia + ((size_t)1 << (s - 1));
  const size_t idx = j << (m - s);
  Xyzz<P> a = load_xyzz<P>(data + ia), b = load_xyzz<P>(data + ib);
  if (idx != 0) {
    Fe<F> w;
    const uint32_t* tw = table + (inverse ? (half - idx) : idx) * 8;
    for (int k = 0; k < 8; k++) w.l[k] = tw[k];
    if (inverse) w = fe_neg<F>(w);                 // w^-idx = -w^(N/2 - idx)
    w = fe_from_mont<F>(w);                        // plain integer for the scalar multiplication
    b = gfft_scalar_mul<C>(tm, b, w.l, glv);
  }
  Xyzz<P> nb = b;
  nb.Y = fe_neg<P>(b.Y);
  gf_add<P>(tm, nb, a);
  gf_add<P>(tm, a, b);
  if (tm.t == 0) {
    store_xyzz<P>(data + ia, a);
    store_xyzz<P>(data + ib, nb);
  }
}

// optional scaling by N^-1 (inverse transform), then normalised projective output
template <class C>
__global__ void __launch_bounds__(128)
k_gfft_store(const XyzzMem<typename C::Fp>* __restrict__ data, const uint32_t* __restrict__ table, int m, int scale, int glv,
             uint32_t* __restrict__ dst) {
  using P = typename C::Fp;
  using F = typename C::Fr;
  constexpr int L = P::L;
  size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 2;   // one team per point
  if (i >= ((size_t)1 << m)) return;
  const Team tm;
  Xyzz<P> a = load_xyzz<P>(data + i);
  if (scale && m > 0) {
    Fe<F> ninv;
    const uint32_t* q = table + ((size_t)1 << (m - 1)) * 8;
    for (int k = 0; k < 8; k++) ninv.l[k] = q[k];
    ninv = fe_from_mont<F>(ninv);
    a = gfft_scalar_mul<C>(tm, a, ninv.l, glv);
  }
  if (tm.t != 0) return;
  uint32_t* o = dst + i * 3 * L;
  Affine<P> aff;
  if (xyzz_to_affine<P>(a, aff)) {
    for (int k = 0; k < L; k++) { o[k] = aff.x.l[k]; o[L + k] = aff.y.l[k]; o[2 * L + k] = P::one(k); }
  } else {
    for (int k = 0; k < L; k++) { o[k] = 0; o[L + k] = P::one(k); o[2 * L + k] = 0; }
  }
}

template <class C>
void gfft_device(cudaStream_t s, int m, const uint32_t* d_gen, const uint32_t* d_src, void* d_work, uint32_t* d_table,
                 uint32_t* d_dst, int inverse, int jac, int glv) {
  using Mem = XyzzMem<typename C::Fp>;
  const size_t N = (size_t)1 << m;
  Mem* data = (Mem*)d_work;
  ntt_build_table<typename C::Fr>(s, d_gen, N >> 1, m, d_table);
  k_gfft_load<C><<<(unsigned)((N + 127) / 128), 128, 0, s>>>(d_src, m, jac, data);
  for (int st = 1; st <= m; st++)
    k_gfft_stage<C><<<(unsigned)((4 * (N >> 1) + 127) / 128), 128, 0, s>>>(data, d_table, m, st, inverse, glv);
  k_gfft_store<C><<<(unsigned)((4 * N + 127) / 128), 128, 0, s>>>(data, d_table, m, inverse, glv, d_dst);
}

template void gfft_device<Bn254>(cudaStream_t, int, const uint32_t*, const uint32_t*, void*, uint32_t*, uint32_t*, int, int, int);
template void gfft_device<Bls12381>(cudaStream_t, int, const uint32_t*, const uint32_t*, void*, uint32_t*, uint32_t*, int, int, int);
template void gfft_device<Bn254G2>(cudaStream_t, int, const uint32_t*, const uint32_t*, void*, uint32_t*, uint32_t*, int, int, int);
template void gfft_device<Bls12381G2>(cudaStream_t, int, const uint32_t*, const uint32_t*, void*, uint32_t*, uint32_t*, int, int, int);

}  // namespace zk
