// Pippenger MSM kernels, shared part: global-memory images of points and the launcher declarations.
// Kernels live in kernels_acc.cuh (recode, accumulate, fixup) and kernels_red.cuh (reduce, tail, sum);
// each is instantiated per curve in its own translation unit so the library builds in parallel.
//
// Replaces the hot loops of the reference's
//   <curve>_G1_proj_MSM_std_coeff_proj_out_variable   lib/cbits/curves/g1/proj/bn128_G1_proj.c:506-586
//   <curve>_G1_proj_MSM_mont_coeff_proj_out           lib/cbits/curves/g1/proj/bn128_G1_proj.c:629-643
// (template: codegen/src/Zikkurat/CodeGen/Curve/MSM.hs:86-166).
//
// Vocabulary: a *segment* is one window of one MSM of a batch (segment = msm * W + window); every
// segment owns NB = 2^(c-1) buckets with weights 1..NB (signed digits).  "pairs" are the
// (bucket key, point index | sign<<31) records, stored segment-major: pair i of segment s at [s*n+i].
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "curves.cuh"
#include "ec.cuh"
#include "recode.cuh"

namespace zk {

// ---- global-memory images (16-byte vector accesses) ---------------------------------------------------
template <class P>
struct alignas(16) XyzzMem {
  uint32_t w[4 * P::L];
};

template <class P>
ZK_D Xyzz<P> load_xyzz(const XyzzMem<P>* p) {
  Xyzz<P> r;
  const uint4* q = reinterpret_cast<const uint4*>(p);
  uint32_t t[4 * P::L];
#pragma unroll
  for (int i = 0; i < P::L; i++) {
    uint4 v = q[i];
    t[4 * i] = v.x; t[4 * i + 1] = v.y; t[4 * i + 2] = v.z; t[4 * i + 3] = v.w;
  }
#pragma unroll
  for (int i = 0; i < P::L; i++) {
    r.X.l[i] = t[i]; r.Y.l[i] = t[P::L + i]; r.ZZ.l[i] = t[2 * P::L + i]; r.ZZZ.l[i] = t[3 * P::L + i];
  }
  return r;
}
template <class P>
ZK_D void store_xyzz(XyzzMem<P>* p, const Xyzz<P>& a) {
  uint32_t t[4 * P::L];
#pragma unroll
  for (int i = 0; i < P::L; i++) {
    t[i] = a.X.l[i]; t[P::L + i] = a.Y.l[i]; t[2 * P::L + i] = a.ZZ.l[i]; t[3 * P::L + i] = a.ZZZ.l[i];
  }
  uint4* q = reinterpret_cast<uint4*>(p);
#pragma unroll
  for (int i = 0; i < P::L; i++) q[i] = make_uint4(t[4 * i], t[4 * i + 1], t[4 * i + 2], t[4 * i + 3]);
}
// affine point i of the caller's array (x || y, 2L 32-bit words, 32-byte aligned records)
template <class P>
ZK_D Affine<P> load_affine(const uint32_t* __restrict__ pts, uint32_t i) {
  Affine<P> r;
  const uint4* q = reinterpret_cast<const uint4*>(pts + (size_t)i * (2 * P::L));
  uint32_t t[2 * P::L];
#pragma unroll
  for (int k = 0; k < P::L / 2; k++) {
    uint4 v = __ldg(q + k);
    t[4 * k] = v.x; t[4 * k + 1] = v.y; t[4 * k + 2] = v.z; t[4 * k + 3] = v.w;
  }
#pragma unroll
  for (int k = 0; k < P::L; k++) { r.x.l[k] = t[k]; r.y.l[k] = t[P::L + k]; }
  return r;
}


constexpr int FIXUP_FAN = 8;  // fan-in per level of the head fix-up tree (kernels_acc.cuh)

enum OutMode : int { OUT_PROJ = 0, OUT_JAC = 1, OUT_AFFINE = 2, OUT_XYZZ = 3 };

// ---- host-side launchers (defined next to their kernels, explicitly instantiated per curve) -------------
template <class C> void launch_recode(cudaStream_t s, const uint64_t* scalars, int nl64, size_t n, int nmsm, int mont,
                                      int nbits, int c, int W, uint32_t* keys, uint32_t* vals);
template <class C> void launch_accumulate(cudaStream_t s, const uint32_t* keys, const uint32_t* vals, const uint32_t* points,
                                          size_t n, int nseg, int chunk, uint32_t chunks_per_seg, uint32_t NB,
                                          XyzzMem<typename C::Fp>* buckets, XyzzMem<typename C::Fp>* heads, uint32_t* head_keys);
template <class C> int accumulate_resident_threads();
template <class C> void launch_fixup_level(cudaStream_t s, const uint32_t* keys_in, const XyzzMem<typename C::Fp>* heads_in,
                                           uint32_t T_in, uint32_t* keys_out, XyzzMem<typename C::Fp>* heads_out, uint32_t T_out,
                                           int nseg, uint32_t NB, XyzzMem<typename C::Fp>* buckets, int last);
template <class C> void launch_reduce_first(cudaStream_t s, const XyzzMem<typename C::Fp>* buckets, int nslices, size_t slice_stride,
                                            size_t total_out, int log_m, XyzzMem<typename C::Fp>* U, XyzzMem<typename C::Fp>* V);
template <class C> void launch_reduce_next(cudaStream_t s, const XyzzMem<typename C::Fp>* Uin, const XyzzMem<typename C::Fp>* Vin,
                                           size_t total_out, int log_m, int log_M, XyzzMem<typename C::Fp>* Uout,
                                           XyzzMem<typename C::Fp>* Vout);
template <class C> void launch_tail(cudaStream_t s, const XyzzMem<typename C::Fp>* Rw, int nmsm, int W, int c, int mode, uint32_t* out);
template <class C> void launch_sum_points(cudaStream_t s, const uint32_t* in, int k, int in_mode, int out_mode, uint32_t* out);
template <class C> void launch_batch_to_affine(cudaStream_t s, const uint32_t* src, size_t n, uint32_t* dst, int jac);
template <class C> void launch_batch_from_affine(cudaStream_t s, const uint32_t* src, size_t n, uint32_t* dst, int jac);
template <class C> void launch_gen_chain(cudaStream_t s, const uint32_t* p0d, unsigned long long start, size_t n, uint32_t* out);

}  // namespace zk
