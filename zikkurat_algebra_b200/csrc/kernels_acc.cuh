// Kernels K1/K2 (recode) and K4 (bucket accumulation + head fix-up) with their launchers.
// See msm_common.cuh for vocabulary.  Reference loops replaced:
//   lib/cbits/curves/g1/proj/bn128_G1_proj.c:520-561 (digit extraction + bucket accumulation)
//   lib/cbits/curves/g1/proj/bn128_G1_proj.c:629-643 (Fr Montgomery -> standard conversion)
#pragma once
#include "ec_team.cuh"
#include <cstdlib>

#include "glv.cuh"
#include "msm_common.cuh"
#include "sort.cuh"

namespace zk {

// ---- K1 + K2: Fr Montgomery -> standard (optional) and signed-digit recoding ---------------------------
// scalars: nmsm*n records of `nl64` 64-bit limbs.  Output pairs are segment-major.
template <class C>
__global__ void __launch_bounds__(256)
k_recode(const uint64_t* __restrict__ scalars, int nl64, size_t n, int nmsm, int mont, int nbits, int c, int W,
         uint2* __restrict__ pairs) {
  size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= n * (size_t)nmsm) return;
  size_t msm = gid / n, i = gid - msm * n;
  Fe<typename C::Fr> k;
  const uint32_t* src = reinterpret_cast<const uint32_t*>(scalars + gid * (size_t)nl64);
  if (nl64 == 4) {
    const uint4* q = reinterpret_cast<const uint4*>(src);
    uint4 a = __ldg(q), b = __ldg(q + 1);
    k.l[0] = a.x; k.l[1] = a.y; k.l[2] = a.z; k.l[3] = a.w;
    k.l[4] = b.x; k.l[5] = b.y; k.l[6] = b.z; k.l[7] = b.w;
  } else {
#pragma unroll
    for (int j = 0; j < 8; j++) k.l[j] = (j < 2 * nl64) ? src[j] : 0u;
  }
  if (mont) k = fe_from_mont<typename C::Fr>(k);
  uint32_t carry = 0;
  size_t seg0 = msm * (size_t)W;
  for (int w = 0; w < W; w++) {
    uint32_t key, neg;
    recode_digit(k.l, nbits, c, w, carry, key, neg);
    pairs[(seg0 + w) * n + i] = make_uint2(key, (uint32_t)i | (neg << 31));
  }
}

// The same with the GLV split (glv.cuh): every scalar becomes two 127-bit halves, k1 over P_i (pair position i of a
// segment) and k2 over phi(P_i) (position n + i, point index n + i of the expanded array); segments hold 2n pairs.
// The sign of a half flips the sign of its digits' points.
template <class C>
__global__ void __launch_bounds__(256)
k_recode_glv(const uint64_t* __restrict__ scalars, int nl64, size_t n, int nmsm, int mont, int c, int W,
             uint2* __restrict__ pairs) {
  using G = typename GlvOf<C>::type;
  size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= n * (size_t)nmsm) return;
  size_t msm = gid / n, i = gid - msm * n;
  Fe<typename C::Fr> k;
  const uint32_t* src = reinterpret_cast<const uint32_t*>(scalars + gid * (size_t)nl64);
  if (nl64 == 4) {
    const uint4* q = reinterpret_cast<const uint4*>(src);
    uint4 a = __ldg(q), b = __ldg(q + 1);
    k.l[0] = a.x; k.l[1] = a.y; k.l[2] = a.z; k.l[3] = a.w;
    k.l[4] = b.x; k.l[5] = b.y; k.l[6] = b.z; k.l[7] = b.w;
  } else {
#pragma unroll
    for (int j = 0; j < 8; j++) k.l[j] = (j < 2 * nl64) ? src[j] : 0u;
  }
  if (mont) k = fe_from_mont<typename C::Fr>(k);
  uint32_t h[2][8];
  bool neg[2];
#pragma unroll
  for (int j = 4; j < 8; j++) h[0][j] = h[1][j] = 0;
  glv_decompose<G>(k.l, h[0], neg[0], h[1], neg[1]);
  const size_t n2 = 2 * n;
  size_t seg0 = msm * (size_t)W;
#pragma unroll
  for (int half = 0; half < 2; half++) {
    uint32_t carry = 0;
    const uint32_t idx = (uint32_t)(half ? n + i : i);
    for (int w = 0; w < W; w++) {
      uint32_t key, dneg;
      recode_digit(h[half], G::BITS, c, w, carry, key, dneg);
      pairs[(seg0 + w) * n2 + idx] = make_uint2(key, idx | ((key != 0 && (dneg != 0) != neg[half]) ? 0x80000000u : 0u));
    }
  }
}

// [P_0 .. P_{n-1} ; phi(P_0) .. phi(P_{n-1})]: the expanded point array of the GLV split.  phi(x, y) = (beta x, y);
// the point at infinity (all-0xFF record) stays what it is.
template <class C>
__global__ void __launch_bounds__(128)
k_glv_points(const uint32_t* __restrict__ src, size_t n, uint32_t* __restrict__ dst) {
  using P = typename C::Fp;
  using G = typename GlvOf<C>::type;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  constexpr int PW = (2 * P::L) / 4;
  const uint4* q = reinterpret_cast<const uint4*>(src + i * (2 * P::L));
  uint4* d0 = reinterpret_cast<uint4*>(dst + i * own_stride<P>());
  uint4* d1 = reinterpret_cast<uint4*>(dst + (n + i) * own_stride<P>());
  uint32_t t[2 * P::L];
#pragma unroll
  for (int kq = 0; kq < PW; kq++) {
    uint4 v = __ldg(q + kq);
    d0[kq] = v;
    t[4 * kq] = v.x; t[4 * kq + 1] = v.y; t[4 * kq + 2] = v.z; t[4 * kq + 3] = v.w;
  }
  Affine<P> p;
#pragma unroll
  for (int kq = 0; kq < P::L; kq++) { p.x.l[kq] = t[kq]; p.y.l[kq] = t[P::L + kq]; }
  if (!affine_is_inf<P>(p)) {
    Fe<P> bx = glv_beta_x<P, G>(p.x);
#pragma unroll
    for (int kq = 0; kq < P::L; kq++) t[kq] = bx.l[kq];
  }
#pragma unroll
  for (int kq = 0; kq < PW; kq++) d1[kq] = make_uint4(t[4 * kq], t[4 * kq + 1], t[4 * kq + 2], t[4 * kq + 3]);
}

// ---- K4: bucket accumulation ---------------------------------------------------------------------------
// Every thread owns `chunk` consecutive SORTED pairs of one segment and folds each run of equal keys
// into one XYZZ sum (xyzz_madd = the IMAD-bound inner operation).  Runs that begin inside the chunk
// are complete from this thread's point of view and are stored straight to their bucket (unique
// writer); the chunk's first run may continue a run of the previous chunk and goes to heads[t]
// instead, to be folded in by k_fixup.  Work per thread is exactly `chunk` insertions whatever the
// scalar distribution, so there is no bucket-size imbalance.
template <class C, bool CALLS, int MINB = (C::Fp::L <= 8 ? 4 : (C::Fp::L <= 12 ? 3 : 2))>
__global__ void __launch_bounds__(128, MINB)
k_accumulate(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ vals,
             const uint32_t* __restrict__ points, int pstride, size_t n, int nseg, int chunk, uint32_t chunks_per_seg,
             uint32_t NB, XyzzMem<typename C::Fp>* __restrict__ buckets,
             XyzzMem<typename C::Fp>* __restrict__ heads, uint32_t* __restrict__ head_keys) {
  using P = typename C::Fp;
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t)nseg * chunks_per_seg) return;
  uint32_t seg = (uint32_t)(t / chunks_per_seg);
  uint32_t j = (uint32_t)(t - (size_t)seg * chunks_per_seg);
  size_t start = (size_t)j * chunk;
  size_t end = start + chunk < n ? start + chunk : n;
  const uint32_t* kp = keys + (size_t)seg * n;
  const uint32_t* vp = vals + (size_t)seg * n;
  XyzzMem<P>* bseg = buckets + (size_t)seg * NB;

  // Software pipeline for the random point gather: while entry e is being added, the point of entry e+1 is
  // already on its way into this thread's shared-memory slot (cp.async, no registers held), and the
  // (key, index) pair of entry e+2 is being read.  Slot layout [buffer][16-byte word][thread]: conflict-free.
  constexpr int PW = (2 * P::L) / 4;                 // 16-byte words per affine point
  __shared__ uint4 stage[2][PW][128];
  auto prefetch = [&](int buf, uint32_t v) {
    const uint4* src = reinterpret_cast<const uint4*>(points + (size_t)(v & 0x7fffffffu) * pstride);
#pragma unroll
    for (int w = 0; w < PW; w++) {
      unsigned dst = (unsigned)__cvta_generic_to_shared(&stage[buf][w][threadIdx.x]);
      asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src + w) : "memory");
    }
  };
  Xyzz<P> acc = xyzz_inf<P>();
  uint32_t cur = 0;      // key of the open run (0 = none)
  bool head_open = true;  // the open run is the chunk's first one
  uint32_t head_key = 0;
  // skip the zero-digit prefix of the segment (digit 0 = no insertion; those pairs sort to the front)
  size_t e = start;
  while (e < end && kp[e] == 0) e++;
  uint32_t key = 0, v = 0, key1 = 0, v1 = 0;
  if (e < end) { key = kp[e]; v = vp[e]; prefetch((int)(e & 1), v); }
  asm volatile("cp.async.commit_group;" ::: "memory");
  if (e + 1 < end) { key1 = kp[e + 1]; v1 = vp[e + 1]; }
  for (; e < end; e++) {
    uint32_t key2 = 0, v2 = 0;
    if (e + 1 < end) prefetch((int)((e + 1) & 1), v1);
    asm volatile("cp.async.commit_group;" ::: "memory");
    if (e + 2 < end) { key2 = kp[e + 2]; v2 = vp[e + 2]; }
    asm volatile("cp.async.wait_group 1;" ::: "memory");   // the copy for entry e (previous group) has landed
    Affine<P> pt;
    {
      uint32_t w32[2 * P::L];
      const int buf = (int)(e & 1);
#pragma unroll
      for (int w = 0; w < PW; w++) {
        uint4 q = stage[buf][w][threadIdx.x];
        w32[4 * w] = q.x; w32[4 * w + 1] = q.y; w32[4 * w + 2] = q.z; w32[4 * w + 3] = q.w;
      }
#pragma unroll
      for (int k = 0; k < P::L; k++) { pt.x.l[k] = w32[k]; pt.y.l[k] = w32[P::L + k]; }
    }
    bool inf = affine_is_inf<P>(pt);
    Fe<P> ny = fe_neg<P>(pt.y);
    if (v >> 31) pt.y = ny;
    if (key != cur) {
      if (cur != 0) {
        if (head_open) { store_xyzz<P>(heads + t, acc); head_key = cur; head_open = false; }
        else store_xyzz<P>(bseg + (cur - 1), acc);
      }
      cur = key;
      acc = inf ? xyzz_inf<P>() : xyzz_from_affine<P>(pt);
    } else if (!inf) {
      xyzz_madd<P, CALLS>(acc, pt);
    }
    key = key1; v = v1; key1 = key2; v1 = v2;
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  if (cur != 0) {
    if (head_open) { store_xyzz<P>(heads + t, acc); head_key = cur; }
    else store_xyzz<P>(bseg + (cur - 1), acc);
  }
  head_keys[t] = head_key;
}

// Fold the per-chunk head sums into their buckets: a segmented tree reduction, FIXUP_FAN entries per
// thread and level.  Input: per segment T_in (key, partial sum) records sorted by key (key 0 = empty).
// Exactly like k_accumulate, a thread's first run may continue the previous thread's and is handed to
// the next level (heads_out); every other run started inside this thread's slice, so this thread is
// the only one of the level to touch that bucket and adds its sum in place.  On the last level
// (one thread per segment) everything goes to the buckets.  Depth is log_FAN(T) whatever the scalar
// distribution (a window holding one single key is the worst case).

// by value for 8 limbs (register ABI), by reference for 12 (see kernels_red.cuh)
template <class P>
__device__ __noinline__ Xyzz<P> xyzz_add_tmv(Team tm, Xyzz<P> a, Xyzz<P> b) { return xyzz_add_team<P>(tm, a, b); }
template <class P>
__device__ __noinline__ void xyzz_add_tmr(const Team& tm, Xyzz<P>& a, const Xyzz<P>& b) { a = xyzz_add_team<P>(tm, a, b); }
template <class P>
__device__ __forceinline__ void xyzz_add_tm(const Team& tm, Xyzz<P>& a, const Xyzz<P>& b) {
  if (P::L <= 8) a = xyzz_add_tmv<P>(tm, a, b); else xyzz_add_tmr<P>(tm, a, b);
}

template <class C>
__global__ void __launch_bounds__(128)
k_fixup_level(const uint32_t* __restrict__ keys_in, const XyzzMem<typename C::Fp>* __restrict__ heads_in, uint32_t T_in,
              uint32_t* __restrict__ keys_out, XyzzMem<typename C::Fp>* __restrict__ heads_out, uint32_t T_out, int nseg,
              uint32_t NB, XyzzMem<typename C::Fp>* __restrict__ buckets, int last) {
  using P = typename C::Fp;
  size_t t = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 2;  // one 4-lane team per output entry
  if (t >= (size_t)nseg * T_out) return;
  Team tm;
  uint32_t seg = (uint32_t)(t / T_out);
  uint32_t j = (uint32_t)(t - (size_t)seg * T_out);
  uint32_t lo = j * FIXUP_FAN, hi = lo + FIXUP_FAN < T_in ? lo + FIXUP_FAN : T_in;
  const uint32_t* kp = keys_in + (size_t)seg * T_in;
  const XyzzMem<P>* hp = heads_in + (size_t)seg * T_in;
  XyzzMem<P>* bseg = buckets + (size_t)seg * NB;
  Xyzz<P> acc = xyzz_inf<P>();
  uint32_t cur = 0, head_key = 0;
  // The slice's first run is handed to the next level only when it really continues the previous
  // slice's last run (same key just before `lo`); otherwise nobody else on this level owns that
  // bucket and the sum is folded in right here, so the upper levels stay empty for ordinary inputs.
  uint32_t prev_key = (lo > 0 && lo < T_in) ? kp[lo - 1] : 0u;
  bool head_open = !last;
  uint32_t e = lo;
  for (;;) {
    uint32_t key = e < hi ? kp[e] : 0xffffffffu;   // sentinel closes the last run
    if (key == 0) { e++; continue; }
    const bool same = key == cur;
    const bool to_head = !same && cur != 0 && head_open && cur == prev_key;
    const bool to_bucket = !same && cur != 0 && !to_head;
    // one inlined addition site: either the next partial sum of the open run, or the bucket's current value
    const XyzzMem<P>* operand = same ? hp + e : (to_bucket ? bseg + (cur - 1) : nullptr);
    if (operand) acc = xyzz_add_team<P>(tm, acc, load_xyzz<P>(operand));
    if (same) { e++; continue; }
    if (to_head) { if (tm.t == 0) store_xyzz<P>(heads_out + t, acc); head_key = cur; }
    if (to_bucket && tm.t == 0) store_xyzz<P>(bseg + (cur - 1), acc);
    if (cur != 0) head_open = false;
    if (e >= hi) break;
    cur = key;
    acc = load_xyzz<P>(hp + e);
    e++;
  }
  if (!last && tm.t == 0) keys_out[t] = head_key;
}

// Code shape of the insertion: 0 = every multiplication inlined (120 KB of SASS for 12 limbs),
// 1 = multiplications out of line (fits the instruction cache).  Measured on B200: 12 limbs run 3-5 %
// faster with 1, 8 limbs 10 % faster with 0 (profiles/r1_notes.md).  Override: $ZKB200_ACC_VARIANT.
template <class C>
inline int accumulate_variant() {
  // read once; initialisation of a function-local static is thread-safe (one host thread per device may get here)
  static const int v = [] { const char* e = getenv("ZKB200_ACC_VARIANT"); return e ? atoi(e) : (C::Fp::L > 8 ? 1 : 0); }();
  return v;
}

template <class C>
void launch_recode(cudaStream_t s, const uint64_t* scalars, int nl64, size_t n, int nmsm, int mont, int nbits, int c, int W,
                   uint2* pairs) {
  size_t tot = (size_t)nmsm * n;
  k_recode<C><<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(scalars, nl64, n, nmsm, mont, nbits, c, W, pairs);
}
template <class C>
void launch_recode_glv(cudaStream_t s, const uint64_t* scalars, int nl64, size_t n, int nmsm, int mont, int c, int W, uint2* pairs) {
  if constexpr (GlvOf<C>::available) {
    size_t tot = (size_t)nmsm * n;
    k_recode_glv<C><<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(scalars, nl64, n, nmsm, mont, c, W, pairs);
  }
}
template <class C>
void launch_glv_points(cudaStream_t s, const uint32_t* src, size_t n, uint32_t* dst) {
  if constexpr (GlvOf<C>::available) k_glv_points<C><<<(unsigned)((n + 127) / 128), 128, 0, s>>>(src, n, dst);
}
template <class C>
void launch_accumulate(cudaStream_t s, const uint32_t* keys, const uint32_t* vals, const uint32_t* points, int pstride, size_t n, int nseg,
                       int chunk, uint32_t chunks_per_seg, uint32_t NB, XyzzMem<typename C::Fp>* buckets,
                       XyzzMem<typename C::Fp>* heads, uint32_t* head_keys) {
  size_t nthreads = (size_t)nseg * chunks_per_seg;
  if (accumulate_variant<C>() == 2)
    k_accumulate<C, true, 4><<<(unsigned)((nthreads + 127) / 128), 128, 0, s>>>(keys, vals, points, pstride, n, nseg, chunk, chunks_per_seg,
                                                                               NB, buckets, heads, head_keys);
  else if (accumulate_variant<C>() == 1)
    k_accumulate<C, true><<<(unsigned)((nthreads + 127) / 128), 128, 0, s>>>(keys, vals, points, pstride, n, nseg, chunk, chunks_per_seg,
                                                                            NB, buckets, heads, head_keys);
  else
    k_accumulate<C, false><<<(unsigned)((nthreads + 127) / 128), 128, 0, s>>>(keys, vals, points, pstride, n, nseg, chunk, chunks_per_seg,
                                                                             NB, buckets, heads, head_keys);
}
template <class C>
int accumulate_resident_threads() {  // threads of k_accumulate<C> that fit on the whole GPU at once
  int dev = 0, sms = 0, blocks = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (accumulate_variant<C>() == 2) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks, k_accumulate<C, true, 4>, 128, 0);
  else if (accumulate_variant<C>() == 1) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks, k_accumulate<C, true>, 128, 0);
  else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks, k_accumulate<C, false>, 128, 0);
  if (blocks < 1) blocks = 1;
  return sms * blocks * 128;
}

template <class C>
void launch_fixup_level(cudaStream_t s, const uint32_t* keys_in, const XyzzMem<typename C::Fp>* heads_in, uint32_t T_in,
                        uint32_t* keys_out, XyzzMem<typename C::Fp>* heads_out, uint32_t T_out, int nseg, uint32_t NB,
                        XyzzMem<typename C::Fp>* buckets, int last) {
  size_t nthreads = (size_t)nseg * T_out * 4;
  k_fixup_level<C><<<(unsigned)((nthreads + 127) / 128), 128, 0, s>>>(keys_in, heads_in, T_in, keys_out, heads_out, T_out, nseg,
                                                                     NB, buckets, last);
}

#define ZK_INSTANTIATE_ACC(C)                                                                                              \
  template int accumulate_resident_threads<C>();                                                                          \
  template void launch_recode<C>(cudaStream_t, const uint64_t*, int, size_t, int, int, int, int, int, uint2*);                  \
  template void launch_recode_glv<C>(cudaStream_t, const uint64_t*, int, size_t, int, int, int, int, uint2*);                \
  template void launch_glv_points<C>(cudaStream_t, const uint32_t*, size_t, uint32_t*);                                    \
  template void launch_accumulate<C>(cudaStream_t, const uint32_t*, const uint32_t*, const uint32_t*, int, size_t, int, int, \
                                     uint32_t, uint32_t, XyzzMem<C::Fp>*, XyzzMem<C::Fp>*, uint32_t*);                      \
  template void launch_fixup_level<C>(cudaStream_t, const uint32_t*, const XyzzMem<C::Fp>*, uint32_t, uint32_t*,           \
                                      XyzzMem<C::Fp>*, uint32_t, int, uint32_t, XyzzMem<C::Fp>*, int);

}  // namespace zk
