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


// Words between consecutive affine records of the library's OWN point arrays (the GLV-expanded points, the affine
// tree's temporaries); the caller's array has 2L-word records, kernels take the stride of `points` as an argument.
// Padding 12-limb records from 96 to 128 bytes (a record then never straddles a 128-byte line, the x coordinate alone is
// one aligned 64-byte DRAM burst) was measured and LOST: BLS12-381 2^20 6.38 -> 6.69 ms -- the arrays grow by a third and
// fall further out of the 126 MB L2, which costs more than the straddling did (profiles/r2_notes.md).
template <class P> __host__ __device__ constexpr int own_stride() { return 2 * P::L; }

constexpr int FIXUP_FAN = 8;  // fan-in per level of the head fix-up tree (kernels_acc.cuh)

enum OutMode : int { OUT_PROJ = 0, OUT_JAC = 1, OUT_AFFINE = 2, OUT_XYZZ = 3 };

// ---- host-side launchers (defined next to their kernels, explicitly instantiated per curve) -------------
// recoding: packed (key, value) pairs (sort.cuh)
template <class C> void launch_recode(cudaStream_t s, const uint64_t* scalars, int nl64, size_t n, int nmsm, int mont,
                                      int nbits, int c, int W, uint2* pairs);
template <class C> void launch_recode_glv(cudaStream_t s, const uint64_t* scalars, int nl64, size_t n, int nmsm, int mont, int c, int W,
                                          uint2* pairs);
template <class C> void launch_glv_points(cudaStream_t s, const uint32_t* src, size_t n, uint32_t* dst);
template <class C> void launch_accumulate(cudaStream_t s, const uint32_t* keys, const uint32_t* vals, const uint32_t* points, int pstride,
                                          size_t n, int nseg, int chunk, uint32_t chunks_per_seg, uint32_t NB,
                                          XyzzMem<typename C::Fp>* buckets, XyzzMem<typename C::Fp>* heads, uint32_t* head_keys);
template <class C> int accumulate_resident_threads();

// ---- affine pre-reduction tree (kernels_aff.cuh) ----
constexpr int AFF_B = 16;         // merges per thread and level (at most; small levels use fewer, down to AFF_B_MIN)
constexpr int AFF_B_MIN = 4;
constexpr int AFF_THREADS = 128;
constexpr int BINV_G = 4;         // elements per thread in the product trees of the batch inversion
struct AffWork {          // device workspaces, sized by aff_sizes()
  uint32_t* tmp;          // temporary affine points
  uint32_t* pre;          // running products, one per merge of the widest level (levels 0, 2, ..)
  uint32_t* pre2;         // the same for the odd levels (half the size)
  uint32_t* binv;         // thread totals + batch_invert workspace
  uint4* st[2];           // block states, ping-pong
  uint32_t* keys_out;     // records for k_accumulate: 2 * ceil(n / 2^R) per segment
  uint32_t* vals_out;
};
struct AffSizes {
  size_t tmp_points, pre_elems, pre2_elems, binv_elems, st0, st1, rec;
  uint32_t nrec;          // records per segment after the last level
};
// Levels of the batch inversion's product tree: big levels are thread-serial (BINV_GS elements per thread, no
// redundant work), small ones use warp scans (32 * BINV_G elements per warp: fewer dependent steps).
constexpr int BINV_GS = 8;
constexpr size_t BINV_SCAN_BELOW = 32768;
struct BinvLevel { size_t T; int kind; };   // kind 0 = serial, 1 = warp scan, 2 = top (one warp, the inversion)
inline int binv_plan(size_t T0, BinvLevel* lv) {
  int n = 0;
  size_t T = T0;
  for (;;) {
    if (T <= 32 * BINV_G) { lv[n++] = {T, 2}; return n; }
    if (T > BINV_SCAN_BELOW) { lv[n++] = {T, 0}; T = (T + BINV_GS - 1) / BINV_GS; }
    else { lv[n++] = {T, 1}; T = (T + 32 * BINV_G - 1) / (32 * BINV_G); }
  }
}
inline size_t binv_workspace_elems(size_t T0) {   // per level: PRE (T), X (T/G + 32, scan levels), next level's elements
  BinvLevel lv[16];
  int n = binv_plan(T0, lv);
  size_t tot = 0;
  for (int i = 0; i < n; i++) tot += lv[i].T + (lv[i].T + BINV_G - 1) / BINV_G + 32 + (i + 1 < n ? lv[i + 1].T + 32 : 0);
  return tot;
}
inline AffSizes aff_sizes(size_t n, int nseg, int R) {
  AffSizes z{};
  size_t nin = n;
  for (int r = 0; r < R; r++) {
    size_t nm = (nin + 1) / 2;
    z.tmp_points += (size_t)nseg * nm;
    if (r == 0) {
      z.pre_elems = (size_t)nseg * nm;
      size_t blocks = ((size_t)nseg * nm + AFF_THREADS * AFF_B_MIN - 1) / (AFF_THREADS * AFF_B_MIN);   // most threads
      z.binv_elems = blocks * AFF_THREADS + binv_workspace_elems(blocks * AFF_THREADS) + 64;
      z.st0 = (size_t)nseg * nm;
    }
    if (r == 1) { z.st1 = (size_t)nseg * nm; z.pre2_elems = (size_t)nseg * nm; }
    nin = nm;
  }
  z.nrec = (uint32_t)(2 * nin);
  z.rec = (size_t)nseg * z.nrec;
  return z;
}
// Lanes of the tree: one or two ranges of segments that run on their own streams, the latency-bound inversion
// chains on high-priority streams, so that one lane's chain runs under the other lane's additions instead of leaving
// the GPU idle.  The caller orders the lane streams against its own stream (events before and after).
struct AffLanes {
  int n;                       // 1 or 2
  int seg0[2], segs[2];        // segment range of each lane
  cudaStream_t big[2], chain[2];
  cudaEvent_t ev_a[2], ev_c[2];
  size_t binv_base0;           // inversion workspace of lane 0 (elements from w.binv)
  size_t binv_stride;          // elements between the two lanes' inversion workspaces
};
// R levels of pairwise affine sums over the sorted pairs of the lanes' segments, then the XYZZ accumulation of the
// surviving records (chunk_rec record slots per thread, cps threads per segment; heads / head_keys indexed by segment
// as for launch_accumulate).  keys / vals / buckets / heads are the arrays of ALL segments.
template <class C> int launch_affine_tree(const AffLanes& ln, const uint32_t* keys, const uint32_t* vals, const uint32_t* points, int pstride,
                                          size_t n, int R, uint32_t NB, XyzzMem<typename C::Fp>* buckets, const AffWork& w,
                                          int chunk_rec, uint32_t cps, XyzzMem<typename C::Fp>* heads, uint32_t* head_keys);
template <class C> void launch_accumulate_rec(cudaStream_t s, const uint32_t* keys, const uint32_t* vals, const uint32_t* points, int pstride,
                                              const uint32_t* tmp_points, size_t n, int nseg, int chunk, uint32_t chunks_per_seg,
                                              uint32_t NB, XyzzMem<typename C::Fp>* buckets, XyzzMem<typename C::Fp>* heads,
                                              uint32_t* head_keys);
template <class C> void launch_fixup_level(cudaStream_t s, const uint32_t* keys_in, const XyzzMem<typename C::Fp>* heads_in,
                                           uint32_t T_in, uint32_t* keys_out, XyzzMem<typename C::Fp>* heads_out, uint32_t T_out,
                                           int nseg, uint32_t NB, XyzzMem<typename C::Fp>* buckets, int last);
template <class C> void launch_reduce_first(cudaStream_t s, const XyzzMem<typename C::Fp>* buckets, int nslices, size_t slice_stride,
                                            size_t total_out, int log_m, XyzzMem<typename C::Fp>* U, XyzzMem<typename C::Fp>* V,
                                            int team);
template <class C> void launch_reduce_next(cudaStream_t s, const XyzzMem<typename C::Fp>* Uin, const XyzzMem<typename C::Fp>* Vin,
                                           size_t total_out, int log_m, XyzzMem<typename C::Fp>* Uout,
                                           XyzzMem<typename C::Fp>* Vout);
template <class C> void launch_tail(cudaStream_t s, const XyzzMem<typename C::Fp>* Rw, int nmsm, int W, int c, int mode, uint32_t* out);
template <class C> void launch_tail_group(cudaStream_t s, const XyzzMem<typename C::Fp>* Rw, int Wg, int c, int extra,
                                          XyzzMem<typename C::Fp>* out);
constexpr int RED2D_MIN_BITS = 6;       // the low-latency reduction needs c - 1 >= this ...
constexpr int RED2D_MAX_WINDOWS = 32;   // ... and at most this many windows per group (teams of k_tail_group_bits)
// low-latency bucket reduction of `ns` consecutive windows (kernels_red.cuh): returns the number of launches
template <class C> int launch_reduce_2d(cudaStream_t s, const XyzzMem<typename C::Fp>* buckets, int nslices, size_t slice_stride, int ns, int c,
                                        int extra, XyzzMem<typename C::Fp>* RC, XyzzMem<typename C::Fp>* T, XyzzMem<typename C::Fp>* out);
template <class C> void launch_sum_points(cudaStream_t s, const uint32_t* in, int k, int in_mode, int out_mode, uint32_t* out);
template <class C> void launch_batch_to_affine(cudaStream_t s, const uint32_t* src, size_t n, uint32_t* dst, int jac);
template <class C> void launch_convert_z(cudaStream_t s, const uint32_t* src, size_t n, uint32_t* z_out);
template <class C> void launch_convert_apply(cudaStream_t s, const uint32_t* src, const uint32_t* zinv, size_t n, uint32_t* dst, int jac);
// batch inversion of T0 non-zero field elements (kernels_aff.cuh): workspace of binv_workspace_elems(T0) elements, result in *inv_out
template <class C> int launch_batch_invert(cudaStream_t s, const uint32_t* E0, size_t T0, uint32_t* ws, uint32_t** inv_out);
template <class C> void launch_batch_from_affine(cudaStream_t s, const uint32_t* src, size_t n, uint32_t* dst, int jac);
template <class C> void launch_gen_chain(cudaStream_t s, const uint32_t* p0d, unsigned long long start, size_t n, uint32_t* out);

}  // namespace zk
