// Affine pre-reduction of the sorted pairs ("batched affine additions") in front of the XYZZ bucket accumulation.
//
// Why: one bucket insertion in XYZZ coordinates costs 8M+2S.  Adding two AFFINE points costs 2M+1S plus one
// field inversion, and n independent inversions cost one inversion plus 3(n-1) multiplications (Montgomery's
// trick): 5M+1S per addition when millions of them share the inversion.  The reference has no counterpart
// (its bucket loop is bn128_G1_proj.c:520-561, one madd_proj_aff per point); the results are the same group
// elements, and the library's output is the canonical affine sum either way.
//
// How: the sorted pairs of every segment are summed pairwise, level by level (aff_plan.cuh explains the block
// bookkeeping that makes this exact for any key distribution).  One level = three steps on the stream:
//   k_aff_prod    every thread walks AFF_B merges, forms the denominators d = x2 - x1 (2y for a doubling) and
//                 stores their running product in front of each merge; thread totals go to E[0]
//   batch_invert  inverts all thread totals: a few tiny kernels (warp-scan product trees, one fe_inv at the top)
//   k_aff_add     walks the same merges backwards, peels 1/d off the inverted total, finishes the additions,
//                 writes the sums to the temporary point array and the merged block states for the next level
// From level 1 on the denominators are already known when the level below finishes: merge m' of level r+1 adds the
// tail of block 2m' and the head of block 2m'+1, which two NEIGHBOURING lanes of level r's k_aff_add have just produced.
// The even lane therefore forms d' = x2 - x1 from registers (one shuffle of the partner's x; operands that are older
// partial sums are fetched) and stores it where level r+1 expects its running products; level r+1's first step then is
// k_aff_prefix, an in-place exclusive prefix product over that array -- coalesced, no gather, no classification --
// instead of k_aff_prod's second gather of every operand's x coordinate.
// After R levels the 2 * ceil(n / 2^R) surviving (key, ref) records per segment go through the ordinary
// k_accumulate / fix-up path; runs that were closed inside a block were already written to their buckets.
// P+P, P+(-P) and infinity operands are classified per merge and keep their exact group-law meaning.
#pragma once
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <utility>
#include <vector>

#include "aff_plan.cuh"
#include "kernels_acc.cuh"
#include "msm_common.cuh"

namespace zk {

#if defined(__CUDACC__)

template <class P>
ZK_D Fe<P> ld_fe(const uint32_t* p) {
  Fe<P> r;
  const uint4* q = reinterpret_cast<const uint4*>(p);
#pragma unroll
  for (int k = 0; k < P::L / 4; k++) {
    uint4 v = q[k];
    r.l[4 * k] = v.x; r.l[4 * k + 1] = v.y; r.l[4 * k + 2] = v.z; r.l[4 * k + 3] = v.w;
  }
  return r;
}
template <class P>
ZK_D void st_fe(uint32_t* p, const Fe<P>& v) {
  uint4* q = reinterpret_cast<uint4*>(p);
#pragma unroll
  for (int k = 0; k < P::L / 4; k++) q[k] = make_uint4(v.l[4 * k], v.l[4 * k + 1], v.l[4 * k + 2], v.l[4 * k + 3]);
}

template <class P>
ZK_D Fe<P> shfl_fe(const Fe<P>& v, int delta, bool up) {
  Fe<P> r;
#pragma unroll
  for (int i = 0; i < P::L; i++) r.l[i] = up ? __shfl_up_sync(0xffffffffu, v.l[i], delta) : __shfl_down_sync(0xffffffffu, v.l[i], delta);
  return r;
}

template <class P>
ZK_D const uint32_t* aff_addr(const uint32_t* points, int pstride, const uint32_t* tmp, uint32_t ref) {
  return (ref & AFF_TEMP) ? tmp + (size_t)(ref & AFF_IDX) * own_stride<P>() : points + (size_t)(ref & AFF_IDX) * pstride;
}
// the point a ref names, sign applied; inf = it is the point at infinity
template <class P>
ZK_D Affine<P> aff_load(const uint32_t* points, int pstride, const uint32_t* tmp, uint32_t ref, bool& inf) {
  const uint32_t* a = aff_addr<P>(points, pstride, tmp, ref);
  Affine<P> p;
  p.x = ld_fe<P>(a);
  p.y = ld_fe<P>(a + P::L);
  inf = affine_is_inf<P>(p);
  if ((ref >> 31) && !inf) p.y = fe_neg<P>(p.y);
  return p;
}
template <class P>
ZK_D void aff_store(uint32_t* tmp, size_t slot, const Affine<P>& p, bool inf) {
  uint32_t* a = tmp + slot * own_stride<P>();
  if (inf) {
    uint4* q = reinterpret_cast<uint4*>(a);
#pragma unroll
    for (int k = 0; k < P::L / 2; k++) q[k] = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
  } else {
    st_fe<P>(a, p.x);
    st_fe<P>(a + P::L, p.y);
  }
}
template <class P>
ZK_D void aff_to_bucket(XyzzMem<P>* b, const Affine<P>& p, bool inf) {
  store_xyzz<P>(b, inf ? xyzz_inf<P>() : xyzz_from_affine<P>(p));
}

// The two blocks of merge i of a segment.  LEVEL0: blocks are the sorted pairs themselves.
struct AffPair {
  uint32_t Lhk, Lhr, Ltk, Ltr, Rhk, Rhr, Rtk, Rtr;
};
template <bool LEVEL0>
ZK_D AffPair aff_read_pair(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ vals, const uint4* __restrict__ st_in,
                           uint32_t nin, uint32_t seg, uint32_t i) {
  AffPair a;
  const size_t base = (size_t)seg * nin + 2 * (size_t)i;
  const bool hasR = 2 * i + 1 < nin;
  if (LEVEL0) {
    a.Lhk = a.Ltk = keys[base];
    a.Lhr = a.Ltr = vals[base];
    a.Rhk = a.Rtk = hasR ? keys[base + 1] : 0u;
    a.Rhr = a.Rtr = hasR ? vals[base + 1] : 0u;
  } else {
    uint4 l = st_in[base];
    uint4 r = hasR ? st_in[base + 1] : make_uint4(0u, 0u, 0u, 0u);
    a.Lhk = l.x; a.Lhr = l.y; a.Ltk = l.z; a.Ltr = l.w;
    a.Rhk = r.x; a.Rhr = r.y; a.Rtk = r.z; a.Rtr = r.w;
  }
  return a;
}

// ---- step 1: denominators and their running products ------------------------------------------------------
// nin = blocks per segment on this level, nm = ceil(nin/2) merges per segment, total = nseg * nm.
// pre[m] = product of the denominators of this thread's earlier merges (written only where merge m adds).
template <class C, bool LEVEL0>
__global__ void __launch_bounds__(AFF_THREADS)
k_aff_prod(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ vals, const uint4* __restrict__ st_in, uint32_t nin,
           uint32_t nm, uint32_t total, const uint32_t* points, int pstride, const uint32_t* tmp, uint32_t* __restrict__ pre,
           uint32_t* __restrict__ tot, int B) {
  using P = typename C::Fp;
  constexpr bool CALLS = (P::L > 8);
  const uint32_t tile = blockIdx.x * (uint32_t)(AFF_THREADS * B);
  Fe<P> run = fe_one<P>();
#pragma unroll 1
  for (int j = 0; j < B; j++) {
    const uint32_t m = tile + (uint32_t)j * AFF_THREADS + threadIdx.x;
    if (m >= total) break;
    const uint32_t seg = m / nm, i = m - seg * nm;
    AffPair a = aff_read_pair<LEVEL0>(keys, vals, st_in, nin, seg, i);
    if (!(a.Ltk == a.Rhk && a.Ltk != 0)) continue;
    // x coordinates decide almost always; the full points are fetched only for the exceptional shapes
    const uint32_t* a1 = aff_addr<P>(points, pstride, tmp, a.Ltr);
    const uint32_t* a2 = aff_addr<P>(points, pstride, tmp, a.Rhr);
    Fe<P> x1 = ld_fe<P>(a1), x2 = ld_fe<P>(a2), d;
    int cls;
    if (x1.l[P::L - 1] == 0xffffffffu || x2.l[P::L - 1] == 0xffffffffu || fe_eq<P>(x1, x2)) {
      bool i1, i2;
      Affine<P> p1 = aff_load<P>(points, pstride, tmp, a.Ltr, i1), p2 = aff_load<P>(points, pstride, tmp, a.Rhr, i2);
      cls = aff_classify<P>(p1, i1, p2, i2, d);
    } else {
      d = fe_sub<P>(x2, x1);
      cls = AFF_ADD;
    }
    if (cls < AFF_ADD) continue;
    st_fe<P>(pre + (size_t)m * P::L, run);
    run = aff_mul<P, CALLS>(run, d);
  }
  st_fe<P>(tot + ((size_t)blockIdx.x * AFF_THREADS + threadIdx.x) * P::L, run);
}

// ---- step 1 for levels whose denominators were handed down by the level below (k_aff_add<NEXT>) ----------------
// pre[m] holds d_m (ONE where merge m adds nothing); it is replaced by the product of this thread's earlier d's.
template <class P>
__global__ void __launch_bounds__(AFF_THREADS)
k_aff_prefix(uint32_t total, uint32_t* __restrict__ pre, uint32_t* __restrict__ tot, int B) {
  constexpr bool CALLS = (P::L > 8);
  const uint32_t tile = blockIdx.x * (uint32_t)(AFF_THREADS * B);
  Fe<P> run = fe_one<P>();
  uint32_t m = tile + threadIdx.x;
  Fe<P> d = m < total ? ld_fe<P>(pre + (size_t)m * P::L) : fe_one<P>();
#pragma unroll 1
  for (int j = 0; j < B; j++, m += AFF_THREADS) {
    if (m >= total) break;
    const uint32_t mn = m + AFF_THREADS;
    Fe<P> dn = (j + 1 < B && mn < total) ? ld_fe<P>(pre + (size_t)mn * P::L) : fe_one<P>();   // next load under this product
    st_fe<P>(pre + (size_t)m * P::L, run);
    run = aff_mul<P, CALLS>(run, d);
    d = dn;
  }
  st_fe<P>(tot + ((size_t)blockIdx.x * AFF_THREADS + threadIdx.x) * P::L, run);
}

// ---- step 3: finish the additions, write sums and merged states ---------------------------------------------
// totinv[t] = inverse of thread t's total.  Sums go to tmp[tmp_off + m].  Output: st_out[m] (uint4 states) or,
// on the LAST level, split (key, ref) records for k_accumulate: 2 per block, head slot 0 when the block is one run.
// The operands of merge j-1 (two points and the stored running product, 5 x field element) travel into this
// thread's shared-memory slot with cp.async while merge j is being computed, exactly like the point gather of
// k_accumulate.  Slot layout [buffer][16-byte word][thread]: conflict-free.
// Build-time experiment switches (tools/build_variant.py): resident blocks per SM asked of ptxas for k_aff_add, and
// whether the stored running product travels through the shared-memory slot too (it costs a fifth of the slot).
// (No default for ZK_AFF_MINB: naming a minimum -- even 1 -- changes ptxas' register allocation; with "1" the 12-limb
// kernels went from 158 to 172 registers, i.e. from three resident blocks to two, and levels 1-2 ran 11 % slower.)
// 12-limb fields: three resident blocks are what the shared-memory slots allow, and the kernels sit right at the
// register count that still permits them (164-172 without a bound, 170 is the limit), so the bound is stated there;
// 0 = unspecified for the other fields.
#ifdef ZK_AFF_MINB
#define ZK_AFF_BOUNDS __launch_bounds__(AFF_THREADS, ZK_AFF_MINB)
#else
#define ZK_AFF_BOUNDS __launch_bounds__(AFF_THREADS, (C::Fp::L == 12 ? 3 : 0))
#endif
#ifndef ZK_AFF_STAGE_PRE
#define ZK_AFF_STAGE_PRE 1
#endif
template <class P>
__host__ __device__ constexpr bool aff_stage_pre() { return ZK_AFF_STAGE_PRE && P::L <= 16; }   // 24 limbs (BLS12-381 Fp2): the points alone fill the slot
template <class P>
constexpr int aff_stage_words() { return 2 * ((2 * P::L) / 4) + (aff_stage_pre<P>() ? P::L / 4 : 0); }
template <class P>
constexpr size_t aff_stage_bytes() { return (size_t)2 * aff_stage_words<P>() * AFF_THREADS * 16; }

// NEXT: also hand the next level its denominators (dnext[m / 2], see the header); needs an even number of merges per
// segment so that the two blocks of a next-level merge sit in neighbouring lanes.
template <class C, bool LEVEL0, bool LAST, bool CALLS, bool NEXT>
__global__ void ZK_AFF_BOUNDS
k_aff_add(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ vals, const uint4* __restrict__ st_in, uint32_t nin,
          uint32_t nm, uint32_t total, const uint32_t* points, int pstride, uint32_t* tmp, uint32_t tmp_off, const uint32_t* __restrict__ pre,
          const uint32_t* __restrict__ totinv, uint4* __restrict__ st_out, uint32_t* __restrict__ keys_out,
          uint32_t* __restrict__ vals_out, uint32_t NB, XyzzMem<typename C::Fp>* __restrict__ buckets, int B,
          uint32_t* __restrict__ dnext) {
  static_assert(!(LAST && NEXT), "the last level has no successor");
  using P = typename C::Fp;
  constexpr bool SPRE = aff_stage_pre<P>();
  constexpr int PW = (2 * P::L) / 4, FW = SPRE ? P::L / 4 : 0, NW = 2 * PW + FW;
  extern __shared__ uint4 aff_stage[];   // [2][NW][AFF_THREADS]
  const uint32_t tile = blockIdx.x * (uint32_t)(AFF_THREADS * B);
  Fe<P> r = ld_fe<P>(totinv + ((size_t)blockIdx.x * AFF_THREADS + threadIdx.x) * P::L);
  // Two steps ahead of the arithmetic: read_pair(j) loads the two blocks of merge j (eight words from global memory),
  // one step ahead: issue(j) starts the copies of its operands -- the addresses come out of those words, so doing both
  // in one step made every warp wait for the loads once per merge (ncu: 6 % of all stall samples on the one ISETP that
  // first touches them).
  auto read_pair = [&](int j, AffPair& a, uint32_t& seg, uint32_t& i) -> bool {
    const uint32_t m = tile + (uint32_t)j * AFF_THREADS + threadIdx.x;
    if (j < 0 || m >= total) return false;
    seg = m / nm;
    i = m - seg * nm;
    a = aff_read_pair<LEVEL0>(keys, vals, st_in, nin, seg, i);
    return true;
  };
  auto issue = [&](int j, int buf, const AffPair& a) {
    if (a.Ltk == a.Rhk && a.Ltk != 0) {
      const uint32_t m = tile + (uint32_t)j * AFF_THREADS + threadIdx.x;
      const uint4* s1 = reinterpret_cast<const uint4*>(aff_addr<P>(points, pstride, tmp, a.Ltr));
      const uint4* s2 = reinterpret_cast<const uint4*>(aff_addr<P>(points, pstride, tmp, a.Rhr));
      const uint4* s3 = reinterpret_cast<const uint4*>(pre + (size_t)m * P::L);
      uint4* dst = aff_stage + (size_t)buf * NW * AFF_THREADS + threadIdx.x;
#pragma unroll
      for (int w = 0; w < NW; w++) {
        const uint4* src = w < PW ? s1 + w : (w < 2 * PW ? s2 + (w - PW) : s3 + (w - 2 * PW));
        unsigned d = (unsigned)__cvta_generic_to_shared(dst + (size_t)w * AFF_THREADS);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src) : "memory");
      }
    }
  };
  auto staged = [&](int buf, int w0, uint32_t* out, int nwords) {
    const uint4* src = aff_stage + (size_t)buf * NW * AFF_THREADS + threadIdx.x;
#pragma unroll
    for (int w = 0; w < nwords; w++) {
      uint4 q = src[(size_t)(w0 + w) * AFF_THREADS];
      out[4 * w] = q.x; out[4 * w + 1] = q.y; out[4 * w + 2] = q.z; out[4 * w + 3] = q.w;
    }
  };
  AffPair a, an, ann;
  uint32_t seg = 0, i = 0, segn = 0, in_ = 0, segnn = 0, inn = 0;
  bool have = read_pair(B - 1, a, seg, i);
  bool have_next = read_pair(B - 2, an, segn, in_);
  if (have) issue(B - 1, (B - 1) & 1, a);
  asm volatile("cp.async.commit_group;" ::: "memory");
#pragma unroll 1
  for (int j = B - 1; j >= 0; j--) {
    const bool have_nn = read_pair(j - 2, ann, segnn, inn);
    if (have_next) issue(j - 1, (j - 1) & 1, an);
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 1;" ::: "memory");
    const uint32_t m = tile + (uint32_t)j * AFF_THREADS + threadIdx.x;
    // what the next level needs from this merge (NEXT): the merged block and, if it was formed, the new sum
    uint32_t o_hk = 0, o_hr = 0, o_tk = 0, o_tr = 0, o_sref = 0;
    bool o_add = false, o_sinf = false;
    Affine<P> o_s;
    if constexpr (NEXT) { o_s.x = fe_zero<P>(); o_s.y = fe_zero<P>(); }
    if (have) {
      AffPlan pl = aff_plan(a.Lhk, a.Lhr, a.Ltk, a.Ltr, a.Rhk, a.Rhr, a.Rtk, a.Rtr);
      XyzzMem<P>* bseg = buckets + (size_t)seg * NB;
      if (pl.add) {
        Affine<P> p1, p2;
        {
          uint32_t w32[2 * P::L];
          staged(j & 1, 0, w32, PW);
#pragma unroll
          for (int k = 0; k < P::L; k++) { p1.x.l[k] = w32[k]; p1.y.l[k] = w32[P::L + k]; }
          staged(j & 1, PW, w32, PW);
#pragma unroll
          for (int k = 0; k < P::L; k++) { p2.x.l[k] = w32[k]; p2.y.l[k] = w32[P::L + k]; }
        }
        const bool i1 = affine_is_inf<P>(p1), i2 = affine_is_inf<P>(p2);
        if ((a.Ltr >> 31) && !i1) p1.y = fe_neg<P>(p1.y);
        if ((a.Rhr >> 31) && !i2) p2.y = fe_neg<P>(p2.y);
        Fe<P> d;
        const int cls = aff_classify<P>(p1, i1, p2, i2, d);
        Affine<P> s = p1;
        bool sinf = false;
        if (cls >= AFF_ADD) {
          Fe<P> pr;
          if constexpr (SPRE) staged(j & 1, 2 * PW, pr.l, FW);
          else pr = ld_fe<P>(pre + (size_t)m * P::L);
          Fe<P> dinv;
          aff_mul_pair<P, CALLS>(r, pr, d, dinv, r);   // 1/d of this merge, and the running inverse without it
          s = aff_finish<P, CALLS>(cls, p1, p2, dinv);
        } else if (cls == AFF_COPY2) {
          s = p2;
        } else if (cls == AFF_INF) {
          sinf = true;
        }
        if (pl.sum_key != 0) aff_to_bucket<P>(bseg + (pl.sum_key - 1), s, sinf);
        else aff_store<P>(tmp, (size_t)tmp_off + m, s, sinf);
        const uint32_t sref = AFF_TEMP | (tmp_off + m);
        if (pl.hr == AFF_SUM) pl.hr = sref;
        if (pl.tr == AFF_SUM) pl.tr = sref;
        if constexpr (NEXT) { o_add = true; o_sref = sref; o_s = s; o_sinf = sinf; }
      } else {
#pragma unroll
        for (int k = 0; k < 2; k++) {
          if (pl.st_key[k] != 0) {
            bool inf;
            Affine<P> p = aff_load<P>(points, pstride, tmp, pl.st_ref[k], inf);
            aff_to_bucket<P>(bseg + (pl.st_key[k] - 1), p, inf);
          }
        }
      }
      if (LAST) {
        const size_t o = ((size_t)seg * nm + i) * 2;
        keys_out[o] = pl.hk == pl.tk ? 0u : pl.hk;
        vals_out[o] = pl.hk == pl.tk ? 0u : pl.hr;
        keys_out[o + 1] = pl.tk;
        vals_out[o + 1] = pl.tk == 0 ? 0u : pl.tr;
      } else {
        st_out[m] = make_uint4(pl.hk, pl.hr, pl.tk, pl.tr);
      }
      if constexpr (NEXT) { o_hk = pl.hk; o_hr = pl.hr; o_tk = pl.tk; o_tr = pl.tr; }
    }
    if constexpr (NEXT) {
      // Lanes 2q and 2q+1 hold the left and the right block of next-level merge m / 2 (tile and nm are even).  The even
      // lane needs the partner's head: key, ref, and the x coordinate when that head is the sum just formed.
      const bool odd = (threadIdx.x & 1) != 0;
      const uint32_t ek = have ? (odd ? o_hk : o_tk) : 0u;      // the run facing the partner
      const uint32_t er = odd ? o_hr : o_tr;
      const bool e_sum = o_add && er == o_sref;
      const uint32_t rk = __shfl_down_sync(0xffffffffu, ek, 1);
      const uint32_t rr = __shfl_down_sync(0xffffffffu, er, 1);
      const int rfl = __shfl_down_sync(0xffffffffu, (e_sum ? 1 : 0) | (o_sinf ? 2 : 0), 1);
      const Fe<P> rx = shfl_fe<P>(o_s.x, 1, false);
      Fe<P> dn = fe_one<P>();
      bool exc = false;
      if (!odd && have && ek == rk && ek != 0) {
        Fe<P> x1 = e_sum ? o_s.x : ld_fe<P>(aff_addr<P>(points, pstride, tmp, er));
        Fe<P> x2 = (rfl & 1) ? rx : ld_fe<P>(aff_addr<P>(points, pstride, tmp, rr));
        const bool i1 = e_sum ? o_sinf : x1.l[P::L - 1] == 0xffffffffu;
        const bool i2 = (rfl & 1) ? (rfl & 2) != 0 : x2.l[P::L - 1] == 0xffffffffu;
        if (i1 || i2 || fe_eq<P>(x1, x2)) exc = true;    // infinity, P + P, P - P: full points below
        else dn = fe_sub<P>(x2, x1);
      }
      if (__any_sync(0xffffffffu, exc)) {
        const Fe<P> ry = shfl_fe<P>(o_s.y, 1, false);
        if (exc) {
          Affine<P> p1, p2;
          bool j1, j2;
          if (e_sum) { p1 = o_s; j1 = o_sinf; } else p1 = aff_load<P>(points, pstride, tmp, er, j1);
          if (rfl & 1) { p2.x = rx; p2.y = ry; j2 = (rfl & 2) != 0; } else p2 = aff_load<P>(points, pstride, tmp, rr, j2);
          Fe<P> d2;
          if (aff_classify<P>(p1, j1, p2, j2, d2) >= AFF_ADD) dn = d2;
        }
      }
      if (!odd && have) st_fe<P>(dnext + (size_t)(m >> 1) * P::L, dn);
    }
    have = have_next; a = an; seg = segn; i = in_;
    have_next = have_nn; an = ann; segn = segnn; in_ = inn;
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}

// ---- step 2: batch inversion of an array of field elements ------------------------------------------------
// One tree level: every thread multiplies BINV_G consecutive elements (exclusive prefixes to PRE), the warp
// multiplies its 32 thread totals with two shuffle scans; X[t] = product of the OTHER lanes' totals, so that
// 1/total_t = X[t] / warp_total.  Warp totals form the next level's elements.
template <class P>
ZK_D void binv_warp_products(const Fe<P>& mine, Fe<P>& others, Fe<P>& all) {
  constexpr bool CALLS = (P::L > 8);
  const int lane = threadIdx.x & 31;
  Fe<P> s = mine;                                   // inclusive prefix product over lanes
#pragma unroll 1
  for (int off = 1; off < 32; off <<= 1) {
    Fe<P> y = shfl_fe<P>(s, off, true);
    Fe<P> z = aff_mul<P, CALLS>(s, y);
    if (lane >= off) s = z;
  }
  Fe<P> q = mine;                                   // inclusive suffix product
#pragma unroll 1
  for (int off = 1; off < 32; off <<= 1) {
    Fe<P> y = shfl_fe<P>(q, off, false);
    Fe<P> z = aff_mul<P, CALLS>(q, y);
    if (lane + off < 32) q = z;
  }
  Fe<P> sx = shfl_fe<P>(s, 1, true), qx = shfl_fe<P>(q, 1, false);
  if (lane == 0) sx = fe_one<P>();
  if (lane == 31) qx = fe_one<P>();
  others = aff_mul<P, CALLS>(sx, qx);
#pragma unroll
  for (int i = 0; i < P::L; i++) all.l[i] = __shfl_sync(0xffffffffu, s.l[i], 31);
}
template <class P, int G = BINV_G>
ZK_D Fe<P> binv_thread_up(const uint32_t* __restrict__ E, uint32_t T, uint32_t* __restrict__ PRE, uint32_t t) {
  constexpr bool CALLS = (P::L > 8);
  Fe<P> run = fe_one<P>();
  if constexpr (G <= 4) {
    // latency-bound levels: all loads first, then the dependent products
    Fe<P> e[G];
#pragma unroll
    for (int k = 0; k < G; k++) {
      const size_t idx = (size_t)t * G + k;
      e[k] = idx < T ? ld_fe<P>(E + idx * P::L) : fe_one<P>();
    }
#pragma unroll
    for (int k = 0; k < G; k++) {
      const size_t idx = (size_t)t * G + k;
      if (idx < T) {
        st_fe<P>(PRE + idx * P::L, run);
        run = aff_mul<P, CALLS>(run, e[k]);
      }
    }
  } else {
#pragma unroll 1
    for (int k = 0; k < G; k++) {
      const size_t idx = (size_t)t * G + k;
      if (idx < T) {
        st_fe<P>(PRE + idx * P::L, run);
        run = aff_mul<P, CALLS>(run, ld_fe<P>(E + idx * P::L));
      }
    }
  }
  return run;
}
template <class P, int G = BINV_G>
ZK_D void binv_thread_down(const uint32_t* __restrict__ E, uint32_t T, uint32_t* __restrict__ PRE, uint32_t t, Fe<P> r) {
  constexpr bool CALLS = (P::L > 8);
  if constexpr (G <= 4) {
    Fe<P> e[G], p[G];
#pragma unroll
    for (int k = 0; k < G; k++) {
      const size_t idx = (size_t)t * G + k;
      if (idx < T) { p[k] = ld_fe<P>(PRE + idx * P::L); e[k] = ld_fe<P>(E + idx * P::L); }
    }
#pragma unroll
    for (int k = G - 1; k >= 0; k--) {
      const size_t idx = (size_t)t * G + k;
      if (idx < T) {
        st_fe<P>(PRE + idx * P::L, aff_mul<P, CALLS>(r, p[k]));
        if (k > 0) r = aff_mul<P, CALLS>(r, e[k]);
      }
    }
  } else {
#pragma unroll 1
    for (int k = G - 1; k >= 0; k--) {
      const size_t idx = (size_t)t * G + k;
      if (idx < T) {
        Fe<P> p = ld_fe<P>(PRE + idx * P::L), e = ld_fe<P>(E + idx * P::L);
        st_fe<P>(PRE + idx * P::L, aff_mul<P, CALLS>(r, p));
        r = aff_mul<P, CALLS>(r, e);
      }
    }
  }
}

template <class P>
__global__ void __launch_bounds__(128) k_binv_up(const uint32_t* __restrict__ E, uint32_t T, uint32_t* __restrict__ PRE,
                                                 uint32_t* __restrict__ X, uint32_t* __restrict__ Eout) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;     // whole warps: no early exit before the shuffles
  Fe<P> mine = binv_thread_up<P>(E, T, PRE, t), others, all;
  binv_warp_products<P>(mine, others, all);
  if ((size_t)t * BINV_G < T) st_fe<P>(X + (size_t)t * P::L, others);
  if ((threadIdx.x & 31) == 31 && (size_t)(t & ~31u) * BINV_G < T) st_fe<P>(Eout + (size_t)(t >> 5) * P::L, all);
}
template <class P>
__global__ void __launch_bounds__(128) k_binv_down(const uint32_t* __restrict__ E, uint32_t T, uint32_t* __restrict__ PRE,
                                                   const uint32_t* __restrict__ X, const uint32_t* __restrict__ INVup) {
  constexpr bool CALLS = (P::L > 8);
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if ((size_t)t * BINV_G >= T) return;
  Fe<P> r = aff_mul<P, CALLS>(ld_fe<P>(INVup + (size_t)(t >> 5) * P::L), ld_fe<P>(X + (size_t)t * P::L));
  binv_thread_down<P>(E, T, PRE, t, r);
}
// thread-serial level: BINV_GS elements per thread, totals are the next level's elements
template <class P>
__global__ void __launch_bounds__(128) k_binv_up_ser(const uint32_t* __restrict__ E, uint32_t T, uint32_t* __restrict__ PRE,
                                                     uint32_t* __restrict__ Eout) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if ((size_t)t * BINV_GS >= T) return;
  st_fe<P>(Eout + (size_t)t * P::L, binv_thread_up<P, BINV_GS>(E, T, PRE, t));
}
template <class P>
__global__ void __launch_bounds__(128) k_binv_down_ser(const uint32_t* __restrict__ E, uint32_t T, uint32_t* __restrict__ PRE,
                                                       const uint32_t* __restrict__ INVup) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if ((size_t)t * BINV_GS >= T) return;
  binv_thread_down<P, BINV_GS>(E, T, PRE, t, ld_fe<P>(INVup + (size_t)t * P::L));
}
// top of the tree: at most 32 * BINV_G elements, one warp, the only field inversion of the whole batch
template <class P>
__global__ void __launch_bounds__(32) k_binv_top(const uint32_t* __restrict__ E, uint32_t T, uint32_t* __restrict__ PRE) {
  constexpr bool CALLS = (P::L > 8);
  const uint32_t t = threadIdx.x;
  Fe<P> mine = binv_thread_up<P>(E, T, PRE, t), others, all;
  binv_warp_products<P>(mine, others, all);
  Fe<P> inv = fe_inv<P>(all);                       // identical data in all lanes: no divergence
  binv_thread_down<P>(E, T, PRE, t, aff_mul<P, CALLS>(inv, others));
}

// Workspace layout of batch_invert (binv_workspace_elems in msm_common.cuh): levels l = 0.. with T[l] elements;
// per level PRE_l (T[l] elements), X_l (ceil(T[l]/G)), E_{l+1} (ceil(T[l]/(32G))).  Result: PRE_0.
template <class P>
int batch_invert(cudaStream_t s, const uint32_t* E0, size_t T0, uint32_t* ws, uint32_t** inv_out) {
  constexpr int L = P::L;
  BinvLevel lv[16];
  const int nl = binv_plan(T0, lv);
  const uint32_t* E[16];
  uint32_t *PRE[16], *X[16];
  int launches = 0;
  E[0] = E0;
  uint32_t* w = ws;
  for (int i = 0; i < nl; i++) {
    const size_t T = lv[i].T;
    PRE[i] = w; w += T * L;
    X[i] = w; w += ((T + BINV_G - 1) / BINV_G + 32) * L;
    if (lv[i].kind == 2) {
      k_binv_top<P><<<1, 32, 0, s>>>(E[i], (uint32_t)T, PRE[i]);
      launches++;
      break;
    }
    uint32_t* En = w; w += (lv[i + 1].T + 32) * L;
    if (lv[i].kind == 0) {
      size_t threads = (T + BINV_GS - 1) / BINV_GS;
      k_binv_up_ser<P><<<(unsigned)((threads + 127) / 128), 128, 0, s>>>(E[i], (uint32_t)T, PRE[i], En);
    } else {
      size_t threads = (((T + BINV_G - 1) / BINV_G) + 31) & ~(size_t)31;
      k_binv_up<P><<<(unsigned)((threads + 127) / 128), 128, 0, s>>>(E[i], (uint32_t)T, PRE[i], X[i], En);
    }
    launches++;
    E[i + 1] = En;
  }
  for (int i = nl - 2; i >= 0; i--) {
    const size_t T = lv[i].T;
    if (lv[i].kind == 0) {
      size_t threads = (T + BINV_GS - 1) / BINV_GS;
      k_binv_down_ser<P><<<(unsigned)((threads + 127) / 128), 128, 0, s>>>(E[i], (uint32_t)T, PRE[i], PRE[i + 1]);
    } else {
      size_t threads = (T + BINV_G - 1) / BINV_G;
      k_binv_down<P><<<(unsigned)((threads + 127) / 128), 128, 0, s>>>(E[i], (uint32_t)T, PRE[i], X[i], PRE[i + 1]);
    }
    launches++;
  }
  *inv_out = PRE[0];
  return launches;
}

// ---- XYZZ accumulation of the surviving records -----------------------------------------------------------------
// k_accumulate (kernels_acc.cuh) over the record list of the last tree level: `chunk` record SLOTS per thread, empty
// slots (key 0) skipped per lane so that every trip of the loop adds a real record; a value with bit 30 set names a
// point of the temporary array.  Same ownership rules: the chunk's first run goes to heads[t] (it may continue the
// previous chunk's last run), every other run is complete for this level and is stored to its bucket.
template <class C, bool CALLS, int MINB = (C::Fp::L <= 8 ? 4 : (C::Fp::L <= 12 ? 3 : 2))>
__global__ void __launch_bounds__(128, MINB)
k_accumulate_rec(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ vals, const uint32_t* __restrict__ points, int pstride,
                 const uint32_t* __restrict__ tmp_points, size_t n, int nseg, int chunk, uint32_t chunks_per_seg, uint32_t NB,
                 XyzzMem<typename C::Fp>* __restrict__ buckets, XyzzMem<typename C::Fp>* __restrict__ heads,
                 uint32_t* __restrict__ head_keys) {
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
  constexpr int PW = (2 * P::L) / 4;
  __shared__ uint4 stage[2][PW][128];
  auto prefetch = [&](int buf, uint32_t v) {
    const uint4* src = reinterpret_cast<const uint4*>(aff_addr<P>(points, pstride, tmp_points, v));
#pragma unroll
    for (int w = 0; w < PW; w++) {
      unsigned dst = (unsigned)__cvta_generic_to_shared(&stage[buf][w][threadIdx.x]);
      asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src + w) : "memory");
    }
  };
  auto skip = [&](size_t e) { while (e < end && kp[e] == 0) e++; return e; };
  Xyzz<P> acc = xyzz_inf<P>();
  uint32_t cur = 0, head_key = 0;
  bool head_open = true;
  size_t e0 = skip(start), e1 = end;
  uint32_t k0 = 0, v0 = 0, k1 = 0, v1 = 0;
  if (e0 < end) { k0 = kp[e0]; v0 = vp[e0]; prefetch(0, v0); e1 = skip(e0 + 1); }
  asm volatile("cp.async.commit_group;" ::: "memory");
  if (e1 < end) { k1 = kp[e1]; v1 = vp[e1]; }
  int b = 0;
  // next record of this thread (point fetched through the cp.async pipeline); false when the chunk is exhausted
  auto step = [&](Affine<P>& pt, bool& inf, uint32_t& key) -> bool {
    if (e0 >= end) return false;
    size_t e2 = end;
    uint32_t k2 = 0, v2 = 0;
    if (e1 < end) { prefetch(b ^ 1, v1); e2 = skip(e1 + 1); }
    asm volatile("cp.async.commit_group;" ::: "memory");
    if (e2 < end) { k2 = kp[e2]; v2 = vp[e2]; }
    asm volatile("cp.async.wait_group 1;" ::: "memory");
    {
      uint32_t w32[2 * P::L];
#pragma unroll
      for (int w = 0; w < PW; w++) {
        uint4 q = stage[b][w][threadIdx.x];
        w32[4 * w] = q.x; w32[4 * w + 1] = q.y; w32[4 * w + 2] = q.z; w32[4 * w + 3] = q.w;
      }
#pragma unroll
      for (int k = 0; k < P::L; k++) { pt.x.l[k] = w32[k]; pt.y.l[k] = w32[P::L + k]; }
    }
    inf = affine_is_inf<P>(pt);
    Fe<P> ny = fe_neg<P>(pt.y);
    if (v0 >> 31) pt.y = ny;
    key = k0;
    e0 = e1; k0 = k1; v0 = v1;
    e1 = e2; k1 = k2; v1 = v2;
    b ^= 1;
    return true;
  };
  for (;;) {
    Affine<P> pt;
    bool inf = false, got;
    uint32_t key = 0;
    // records that open a new run are cheap (store the finished sum, restart from the point): a lane works
    // through them here, so that the warp meets again at the expensive addition below
    while ((got = step(pt, inf, key)) && key != cur) {
      if (cur != 0) {
        if (head_open) { store_xyzz<P>(heads + t, acc); head_key = cur; head_open = false; }
        else store_xyzz<P>(bseg + (cur - 1), acc);
      }
      cur = key;
      acc = inf ? xyzz_inf<P>() : xyzz_from_affine<P>(pt);
    }
    if (!got) break;
    if (!inf) xyzz_madd<P, CALLS>(acc, pt);
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  if (cur != 0) {
    if (head_open) { store_xyzz<P>(heads + t, acc); head_key = cur; }
    else store_xyzz<P>(bseg + (cur - 1), acc);
  }
  head_keys[t] = head_key;
}

template <class C>
void launch_accumulate_rec(cudaStream_t s, const uint32_t* keys, const uint32_t* vals, const uint32_t* points, int pstride,
                           const uint32_t* tmp_points, size_t n, int nseg, int chunk, uint32_t chunks_per_seg, uint32_t NB,
                           XyzzMem<typename C::Fp>* buckets, XyzzMem<typename C::Fp>* heads, uint32_t* head_keys) {
  size_t nthreads = (size_t)nseg * chunks_per_seg;
  k_accumulate_rec<C, (C::Fp::L > 8)><<<(unsigned)((nthreads + 127) / 128), 128, 0, s>>>(
      keys, vals, points, pstride, tmp_points, n, nseg, chunk, chunks_per_seg, NB, buckets, heads, head_keys);
}

// ---- host helpers -------------------------------------------------------------------------------------------------
#define ZK_AFF_CK(x)                                                                                            \
  do {                                                                                                          \
    cudaError_t e__ = (x);                                                                                      \
    if (e__ != cudaSuccess) {                                                                                   \
      fprintf(stderr, "[zkmsm_b200] fatal: %s failed at %s:%d: %s\n", #x, __FILE__, __LINE__, cudaGetErrorString(e__)); \
      abort();                                                                                                  \
    }                                                                                                           \
  } while (0)
// The opt-in for more than 48 KB of dynamic shared memory is a per-device function attribute: set it ONCE per kernel
// instantiation and device (bit mask), not before every launch.
template <class K>
inline void aff_allow_smem(K kernel, size_t smem) {
  static std::mutex mu;                                            // K is the same function TYPE for all instantiations with
  static std::vector<std::pair<const void*, unsigned>> seen;      // one signature, so the mask is keyed by the function pointer
  int dev = 0;
  ZK_AFF_CK(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lk(mu);
  for (auto& e : seen)
    if (e.first == (const void*)kernel) {
      if (e.second & (1u << dev)) return;
      ZK_AFF_CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      e.second |= 1u << dev;
      return;
    }
  ZK_AFF_CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  seen.emplace_back((const void*)kernel, 1u << dev);
}

// ---- host driver: R levels over sorted pairs, then the records ------------------------------------------------------
template <class C>
int launch_affine_tree(const AffLanes& ln, const uint32_t* keys, const uint32_t* vals, const uint32_t* points, int pstride, size_t n, int R,
                       uint32_t NB, XyzzMem<typename C::Fp>* buckets, const AffWork& w, int chunk_rec, uint32_t cps,
                       XyzzMem<typename C::Fp>* heads, uint32_t* head_keys) {
  using P = typename C::Fp;
  int launches = 0;
  // code shape of the additions: multiplications inlined (1) or out of line (0); $ZKB200_AFF_INLINE
  static const int inl = [] { const char* e = getenv("ZKB200_AFF_INLINE"); return e ? atoi(e) : (P::L > 8 ? 0 : 1); }();
  const size_t smem = aff_stage_bytes<P>();
  const int G = ln.n;
  const cudaStream_t* big = ln.big;
  const cudaStream_t* chain = ln.chain;
  const int* seg0 = ln.seg0;
  const int* segs = ln.segs;
  size_t tmp_per_seg = 0;
  {
    size_t nin = n;
    for (int r = 0; r < R; r++) { nin = (nin + 1) / 2; tmp_per_seg += nin; }
  }
  const size_t tmp_base[2] = {(size_t)seg0[0] * tmp_per_seg, (size_t)seg0[1] * tmp_per_seg};
  const size_t binv_base[2] = {ln.binv_base0, ln.binv_base0 + ln.binv_stride};
  // per level: blocks per segment going in, merges per segment
  uint32_t nin_l[16], nm_l[16];
  {
    uint32_t nin = (uint32_t)n;
    for (int r = 0; r < R; r++) { nin_l[r] = nin; nm_l[r] = (nin + 1) / 2; nin = nm_l[r]; }
  }
  // A segment's block states live at a FIXED place of each ping-pong buffer (stride = the widest level that uses
  // the buffer): lanes are not in lock step, so a level-dependent offset would let one lane's narrow level land in
  // another lane's wide one.
  const size_t st_stride[2] = {nm_l[0], R > 1 ? nm_l[1] : 0};
  size_t lvl_off[2] = {tmp_base[0], tmp_base[1]};   // where the current level's sums start in the temporary array
  // merges per thread: AFF_B (measured: fewer merges per thread = more threads = more inversion-chain work, a loss
  // even for small levels); $ZKB200_AFF_B overrides, down to AFF_B_MIN
  static const int forced_B = [] { const char* e = getenv("ZKB200_AFF_B"); return e ? atoi(e) : 0; }();
  auto pick_B = [&](uint32_t) -> int {
    if (forced_B > 0) return forced_B > 128 ? 128 : (forced_B < AFF_B_MIN ? AFF_B_MIN : forced_B);
    return AFF_B;
  };
  uint32_t* inv[2] = {nullptr, nullptr};
  // Running products of level r live in w.pre (even r) or w.pre2 (odd r): level r's additions read theirs while they
  // write the next level's denominators into the other array.
  // Measured (profiles/r2_notes.md section 9): the prefix kernels save 0.13 ms of serialised kernel time per BLS12-381
  // 2^20 MSM but the additions grow by 0.22 ms (shuffles, a second classification, a fetch for every operand that is not
  // the new sum), so the hand-over is opt-in ($ZKB200_AFF_NEXT=1; read per call so that the tests can switch it).
  const bool next_on = [] { const char* e = getenv("ZKB200_AFF_NEXT"); return e ? atoi(e) != 0 : false; }();
  bool next_l[16];   // level r hands level r+1 its denominators
  for (int r = 0; r < R; r++) next_l[r] = next_on && r + 1 < R && (nm_l[r] % 2 == 0);
  auto pre_of = [&](int g, int r) -> uint32_t* {
    return (r & 1) ? w.pre2 + (size_t)seg0[g] * nm_l[1] * P::L : w.pre + (size_t)seg0[g] * nm_l[0] * P::L;
  };
  // step 1 + 2 of level r for group g: running products on the group's stream, inversion chain on its chain stream
  auto do_prod = [&](int g, int r) {
    const uint32_t nin = nin_l[r], nm = nm_l[r], total = (uint32_t)segs[g] * nm;
    const int B = pick_B(total);
    const unsigned blocks = (total + AFF_THREADS * B - 1) / (AFF_THREADS * B);
    const uint32_t* kg = keys + (size_t)seg0[g] * n;
    const uint32_t* vg = vals + (size_t)seg0[g] * n;
    const uint4* st_in = r == 0 ? nullptr : w.st[(r - 1) & 1] + (size_t)seg0[g] * st_stride[(r - 1) & 1];
    uint32_t* pre = pre_of(g, r);
    uint32_t* tot = w.binv + binv_base[g] * P::L;
    const size_t T0 = (size_t)blocks * AFF_THREADS;
    if (r == 0) k_aff_prod<C, true><<<blocks, AFF_THREADS, 0, big[g]>>>(kg, vg, st_in, nin, nm, total, points, pstride, w.tmp, pre, tot, B);
    else if (next_l[r - 1]) k_aff_prefix<P><<<blocks, AFF_THREADS, 0, big[g]>>>(total, pre, tot, B);
    else k_aff_prod<C, false><<<blocks, AFF_THREADS, 0, big[g]>>>(kg, vg, st_in, nin, nm, total, points, pstride, w.tmp, pre, tot, B);
    launches++;
    if (big[g] != chain[g]) { ZK_AFF_CK(cudaEventRecord(ln.ev_a[g], big[g])); ZK_AFF_CK(cudaStreamWaitEvent(chain[g], ln.ev_a[g], 0)); }
    launches += batch_invert<P>(chain[g], tot, T0, tot + T0 * P::L, &inv[g]);
    if (big[g] != chain[g]) { ZK_AFF_CK(cudaEventRecord(ln.ev_c[g], chain[g])); ZK_AFF_CK(cudaStreamWaitEvent(big[g], ln.ev_c[g], 0)); }
  };
  // step 3 of level r for group g
  auto do_add = [&](int g, int r) {
    const uint32_t nin = nin_l[r], nm = nm_l[r], total = (uint32_t)segs[g] * nm;
    const int B = pick_B(total);
    const unsigned blocks = (total + AFF_THREADS * B - 1) / (AFF_THREADS * B);
    const bool last = r == R - 1;
    const uint32_t* kg = keys + (size_t)seg0[g] * n;
    const uint32_t* vg = vals + (size_t)seg0[g] * n;
    const uint4* st_in = r == 0 ? nullptr : w.st[(r - 1) & 1] + (size_t)seg0[g] * st_stride[(r - 1) & 1];
    uint4* st_out = w.st[r & 1] + (size_t)seg0[g] * st_stride[r & 1];
    uint32_t* pre = pre_of(g, r);
    uint32_t* dnext = next_l[r] ? pre_of(g, r + 1) : nullptr;
    uint32_t* ko = w.keys_out + (size_t)seg0[g] * 2 * nm;
    uint32_t* vo = w.vals_out + (size_t)seg0[g] * 2 * nm;
    XyzzMem<P>* bg = buckets + (size_t)seg0[g] * NB;
    const uint32_t tmp_off = (uint32_t)lvl_off[g];
#define ZK_AFF_ADD(L0, LA, NX)                                                                                                \
  do {                                                                                                                        \
    if (inl) {                                                                                                                \
      aff_allow_smem(k_aff_add<C, L0, LA, false, NX>, smem);                                                                  \
      k_aff_add<C, L0, LA, false, NX><<<blocks, AFF_THREADS, smem, big[g]>>>(kg, vg, st_in, nin, nm, total, points, pstride, w.tmp,   \
                                                                             tmp_off, pre, inv[g], st_out, ko, vo, NB, bg, B, dnext); \
    } else {                                                                                                                  \
      aff_allow_smem(k_aff_add<C, L0, LA, true, NX>, smem);                                                                   \
      k_aff_add<C, L0, LA, true, NX><<<blocks, AFF_THREADS, smem, big[g]>>>(kg, vg, st_in, nin, nm, total, points, pstride, w.tmp,    \
                                                                            tmp_off, pre, inv[g], st_out, ko, vo, NB, bg, B, dnext);  \
    }                                                                                                                         \
  } while (0)
    if (r == 0 && last) ZK_AFF_ADD(true, true, false);
    else if (r == 0 && next_l[r]) ZK_AFF_ADD(true, false, true);
    else if (r == 0) ZK_AFF_ADD(true, false, false);
    else if (last) ZK_AFF_ADD(false, true, false);
    else if (next_l[r]) ZK_AFF_ADD(false, false, true);
    else ZK_AFF_ADD(false, false, false);
#undef ZK_AFF_ADD
    launches++;
    lvl_off[g] += total;
  };
  const uint32_t nrec = 2 * nm_l[R - 1];   // surviving record slots per segment
  auto do_records = [&](int g) {
    launch_accumulate_rec<C>(big[g], w.keys_out + (size_t)seg0[g] * nrec, w.vals_out + (size_t)seg0[g] * nrec, points, pstride, w.tmp, nrec,
                             segs[g], chunk_rec, cps, NB, buckets + (size_t)seg0[g] * NB, heads + (size_t)seg0[g] * cps,
                             head_keys + (size_t)seg0[g] * cps);
    launches++;
  };
  // Enqueue order = execution order of the big kernels (equal-priority streams drain first come first served):
  // a group's next running products follow its additions immediately, so that its inversion chain runs under the
  // OTHER group's additions.  Only the very first chain has just the other group's running products to hide behind.
  for (int g = 0; g < G; g++) do_prod(g, 0);
  for (int r = 0; r < R; r++) {
    for (int g = 0; g < G; g++) {
      do_add(g, r);
      if (r + 1 < R) do_prod(g, r + 1);
      else do_records(g);
    }
  }
  return launches;
}

template <class C>
int launch_batch_invert(cudaStream_t s, const uint32_t* E0, size_t T0, uint32_t* ws, uint32_t** inv_out) {
  return batch_invert<typename C::Fp>(s, E0, T0, ws, inv_out);
}

#define ZK_INSTANTIATE_AFF(C)                                                                                             \
  template int launch_batch_invert<C>(cudaStream_t, const uint32_t*, size_t, uint32_t*, uint32_t**);                      \
  template int launch_affine_tree<C>(const AffLanes&, const uint32_t*, const uint32_t*, const uint32_t*, int, size_t, int, \
                                     uint32_t, XyzzMem<C::Fp>*, const AffWork&, int, uint32_t, XyzzMem<C::Fp>*,           \
                                     uint32_t*);                                                                          \
  template void launch_accumulate_rec<C>(cudaStream_t, const uint32_t*, const uint32_t*, const uint32_t*, int, const uint32_t*, \
                                         size_t, int, int, uint32_t, uint32_t, XyzzMem<C::Fp>*, XyzzMem<C::Fp>*, uint32_t*);

#endif  // __CUDACC__

}  // namespace zk
