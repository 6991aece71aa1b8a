// Element-wise self-test kernels: the field and group primitives exactly as ptxas compiled them (the same
// headers, the same inline-PTX carry chains the MSM kernels use), applied to arrays of operands so that a test can
// compare every result with the reference C (tests/test_device_primitives.py).
//
// Mirrors the reference's fast-vs-reference field tests (test/src/ZK/Test/Field/AgainstRef.hs:25-60) and the group-law
// edge cases of test/src/ZK/Test/Curve/Properties.hs:425-483 for
//   <curve>_Fp_mont_mul/_sqr/_add/_sub/_neg/_inv   lib/cbits/curves/fields/mont/bn128_Fp_mont.c:44-109,177-204
//   <curve>_Fr_mont_to_std                         lib/cbits/curves/fields/mont/bn128_Fr_mont.c:330-335
//   <curve>_G1_proj_madd_proj_aff / _add / _dbl    lib/cbits/curves/g1/proj/bn128_G1_proj.c:230-373
// Not on the product path: nothing in the MSM calls these kernels.
#include <cuda_runtime.h>

#include "msm_common.cuh"
#include "selftest.cuh"

namespace zk {

template <class P>
__device__ __forceinline__ Fe<P> st_ld(const uint32_t* p, size_t i) {
  Fe<P> r;
#pragma unroll
  for (int k = 0; k < P::L; k++) r.l[k] = p[i * P::L + k];
  return r;
}
template <class P>
__device__ __forceinline__ void st_st(uint32_t* p, size_t i, const Fe<P>& v) {
#pragma unroll
  for (int k = 0; k < P::L; k++) p[i * P::L + k] = v.l[k];
}

// out[i] = op(a[i], b[i], c[i], d[i]) over the field P
template <class P>
__global__ void __launch_bounds__(128) k_selftest_field(int op, size_t n, const uint32_t* __restrict__ a, const uint32_t* __restrict__ b,
                                                        const uint32_t* __restrict__ c, const uint32_t* __restrict__ d,
                                                        uint32_t* __restrict__ out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fe<P> x = st_ld<P>(a, i), y = b ? st_ld<P>(b, i) : x, r;
  switch (op) {
    case ZKT_MUL: r = fe_mul<P>(x, y); break;
    case ZKT_SQR: r = fe_sqr<P>(x); break;
    case ZKT_MUL2:
      if constexpr (P::THREE_MOD_FITS) r = fe_mul2<P>(x, y, st_ld<P>(c, i), st_ld<P>(d, i));
      else r = fe_add<P>(fe_mul<P>(x, y), fe_mul<P>(st_ld<P>(c, i), st_ld<P>(d, i)));
      break;
    case ZKT_ADD: r = fe_add<P>(x, y); break;
    case ZKT_SUB: r = fe_sub<P>(x, y); break;
    case ZKT_NEG: r = fe_neg<P>(x); break;
    case ZKT_INV: r = fe_is_zero<P>(x) ? x : fe_inv<P>(x); break;
    case ZKT_MUL_CALL: r = fe_mul_call<P>(x, y); break;
    case ZKT_SQR_CALL: r = fe_sqr_call<P>(x); break;
    case ZKT_MUL2_CALL:
      if constexpr (P::THREE_MOD_FITS) r = fe_mul2_call<P>(x, y, st_ld<P>(c, i), st_ld<P>(d, i));
      else r = fe_add<P>(fe_mul<P>(x, y), fe_mul<P>(st_ld<P>(c, i), st_ld<P>(d, i)));
      break;
    case ZKT_DBL: r = fe_dbl<P>(x); break;
    case ZKT_FROM_MONT: r = fe_from_mont<P>(x); break;
    case ZKT_PAIR_FIRST:
    case ZKT_PAIR_SECOND: {
      FePair<P> pr = fe_mul_pair_call<P>(x, y, st_ld<P>(c, i));
      r = op == ZKT_PAIR_FIRST ? pr.u : pr.v;
      break;
    }
    default: r = fe_zero<P>(); break;
  }
  st_st<P>(out, i, r);
}

// Group operations on XYZZ representatives with NON-trivial denominators: operand k is the affine point p_k lifted
// with the scale z_k: (x z^2, y z^3, z^2, z^3).  An all-0xFF record is the point at infinity.  Result: canonical
// affine bytes (all 0xFF for infinity), which is what the reference's own operation + to_affine gives.
template <class P>
__device__ __forceinline__ Xyzz<P> st_lift(const Affine<P>& p, const Fe<P>& z) {
  if (affine_is_inf<P>(p)) return xyzz_inf<P>();
  Xyzz<P> r;
  r.ZZ = fe_sqr<P>(z);
  r.ZZZ = fe_mul<P>(r.ZZ, z);
  r.X = fe_mul<P>(p.x, r.ZZ);
  r.Y = fe_mul<P>(p.y, r.ZZZ);
  return r;
}
template <class C>
__global__ void __launch_bounds__(64) k_selftest_group(int op, size_t n, const uint32_t* __restrict__ p1, const uint32_t* __restrict__ z1,
                                                       const uint32_t* __restrict__ p2, const uint32_t* __restrict__ z2,
                                                       uint32_t* __restrict__ out) {
  using P = typename C::Fp;
  constexpr int L = P::L;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Affine<P> a, b;
  a.x = st_ld<P>(p1, 2 * i); a.y = st_ld<P>(p1, 2 * i + 1);
  b.x = st_ld<P>(p2, 2 * i); b.y = st_ld<P>(p2, 2 * i + 1);
  Xyzz<P> A = st_lift<P>(a, st_ld<P>(z1, i)), R;
  switch (op) {
    case ZKT_G_MADD:        // the bucket insertion: XYZZ += affine (the caller skips infinity operands, as the MSM does)
      R = A;
      if (!affine_is_inf<P>(b)) xyzz_madd<P, false>(R, b);
      break;
    case ZKT_G_MADD_CALLS:
      R = A;
      if (!affine_is_inf<P>(b)) xyzz_madd<P, true>(R, b);
      break;
    case ZKT_G_ADD: R = xyzz_add<P>(A, st_lift<P>(b, st_ld<P>(z2, i))); break;
    case ZKT_G_ADD_CALLS: R = xyzz_add_calls<P>(A, st_lift<P>(b, st_ld<P>(z2, i))); break;
    case ZKT_G_DBL: R = xyzz_dbl<P>(A); break;
    case ZKT_G_DBL_AFFINE: R = affine_is_inf<P>(a) ? xyzz_inf<P>() : xyzz_dbl_affine<P>(a); break;
    default: R = xyzz_inf<P>(); break;
  }
  Affine<P> o;
  if (xyzz_to_affine<P>(R, o)) { st_st<P>(out, 2 * i, o.x); st_st<P>(out, 2 * i + 1, o.y); }
  else { for (int k = 0; k < 2 * L; k++) out[i * 2 * L + k] = 0xffffffffu; }
}

template <class P>
void launch_selftest_field(cudaStream_t s, int op, size_t n, const uint32_t* a, const uint32_t* b, const uint32_t* c, const uint32_t* d,
                           uint32_t* out) {
  if (n) k_selftest_field<P><<<(unsigned)((n + 127) / 128), 128, 0, s>>>(op, n, a, b, c, d, out);
}
template <class C>
void launch_selftest_group(cudaStream_t s, int op, size_t n, const uint32_t* p1, const uint32_t* z1, const uint32_t* p2, const uint32_t* z2,
                           uint32_t* out) {
  if (n) k_selftest_group<C><<<(unsigned)((n + 63) / 64), 64, 0, s>>>(op, n, p1, z1, p2, z2, out);
}

template void launch_selftest_field<Bn254Fp>(cudaStream_t, int, size_t, const uint32_t*, const uint32_t*, const uint32_t*, const uint32_t*, uint32_t*);
template void launch_selftest_field<Bn254Fr>(cudaStream_t, int, size_t, const uint32_t*, const uint32_t*, const uint32_t*, const uint32_t*, uint32_t*);
template void launch_selftest_field<Bls12381Fp>(cudaStream_t, int, size_t, const uint32_t*, const uint32_t*, const uint32_t*, const uint32_t*, uint32_t*);
template void launch_selftest_field<Bls12381Fr>(cudaStream_t, int, size_t, const uint32_t*, const uint32_t*, const uint32_t*, const uint32_t*, uint32_t*);
template void launch_selftest_group<Bn254>(cudaStream_t, int, size_t, const uint32_t*, const uint32_t*, const uint32_t*, const uint32_t*, uint32_t*);
template void launch_selftest_group<Bls12381>(cudaStream_t, int, size_t, const uint32_t*, const uint32_t*, const uint32_t*, const uint32_t*, uint32_t*);

}  // namespace zk
