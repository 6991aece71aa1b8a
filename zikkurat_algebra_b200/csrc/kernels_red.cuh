// Kernels K5 (bucket reduction by levels), K6/K7 (window combination, output conversion) and K8
// (sum of partial results) with their launchers.  Reference loops replaced:
//   lib/cbits/curves/g1/proj/bn128_G1_proj.c:565-582 (running sums, Horner over windows)
//   lib/cbits/curves/g1/proj/bn128_G1_proj.c:132-144 (to_affine)
#pragma once
#include "ec_team.cuh"
#include "msm_common.cuh"
#include "red_plan.cuh"

namespace zk {

// Out-of-line group operations for the latency-bound kernels (reduce_next, tail, sum): keeps their code
// size and compile time down; the throughput kernels (accumulate, reduce_first) inline everything.
template <class P>
__device__ __noinline__ void xyzz_add_nc(Xyzz<P>& a, const Xyzz<P>& b) { a = xyzz_add<P>(a, b); }
template <class P>
__device__ __noinline__ void xyzz_dbl_nc(Xyzz<P>& a) { a = xyzz_dbl<P>(a); }
// 4-lane team versions (ec_team.cuh), also out of line
// Calling convention matters for these latency-bound kernels: 8-limb points fit the register ABI when
// passed by value (no stack traffic, measured 15-30 % faster); 12-limb points do not fit and by-value
// would add copies through the local stack, so they are passed by reference (measured faster).
template <class P>
__device__ __noinline__ Xyzz<P> xyzz_add_tmv(Team tm, Xyzz<P> a, Xyzz<P> b) { return xyzz_add_team<P>(tm, a, b); }
template <class P>
__device__ __noinline__ Xyzz<P> xyzz_dbl_tmv(Team tm, Xyzz<P> a) { return xyzz_dbl_team<P>(tm, a); }
template <class P>
__device__ __noinline__ void xyzz_add_tmr(const Team& tm, Xyzz<P>& a, const Xyzz<P>& b) { a = xyzz_add_team<P>(tm, a, b); }
template <class P>
__device__ __noinline__ void xyzz_dbl_tmr(const Team& tm, Xyzz<P>& a) { a = xyzz_dbl_team<P>(tm, a); }
template <class P>
__device__ __forceinline__ void xyzz_add_tm(const Team& tm, Xyzz<P>& a, const Xyzz<P>& b) {
  if (P::L <= 8) a = xyzz_add_tmv<P>(tm, a, b); else xyzz_add_tmr<P>(tm, a, b);
}
template <class P>
__device__ __forceinline__ void xyzz_dbl_tm(const Team& tm, Xyzz<P>& a) {
  if (P::L <= 8) a = xyzz_dbl_tmv<P>(tm, a); else xyzz_dbl_tmr<P>(tm, a);
}

// ---- K5: bucket reduction  sum_b (b+1) * B[b]  by levels ------------------------------------------------
// Invariant after every level:  R_seg = sum_t ( U[t] + t * Vs[t] ),  t = 0..S-1, where Vs[t] is the plain sum of the
// entry's buckets ALREADY multiplied by the entry's index stride (Vs = M * V): a level then needs only log2(m)
// doublings of its V output, whatever its height, instead of log2(M) doublings of the weighted sum.
// Level 1 reads the buckets (weights t+1) with the classic running sum over m buckets.
// With K input slices (H2D overlap, see run_msm) there are K bucket arrays `slice_stride` apart; their
// sum is taken on the fly: run += B_0[i] + ... + B_{K-1}[i] through the same addition site.
// 12-limb fields: out-of-line multiplications (250 -> ~170 registers, 3 CTAs per SM); 8 limbs: inlined
#define RF_ADD(x, y) (P::L > 8 ? xyzz_add_calls<P>((x), (y)) : xyzz_add<P>((x), (y)))
template <class C>
__global__ void __launch_bounds__(128)
k_reduce_first(const XyzzMem<typename C::Fp>* __restrict__ buckets, int nslices, size_t slice_stride, size_t total_out,
               int log_m, XyzzMem<typename C::Fp>* __restrict__ U, XyzzMem<typename C::Fp>* __restrict__ V) {
  using P = typename C::Fp;
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total_out) return;
  const XyzzMem<P>* b = buckets + (t << log_m);
  Xyzz<P> run = xyzz_inf<P>(), acc = xyzz_inf<P>();
  for (int i = (1 << log_m) - 1; i >= 0; i--) {
#pragma unroll 1
    for (int k = 0; k < nslices; k++) run = RF_ADD(run, load_xyzz<P>(b + (size_t)k * slice_stride + i));
    acc = RF_ADD(acc, run);
  }
  store_xyzz<P>(U + t, acc);
#pragma unroll 1
  for (int d = 0; d < log_m; d++) xyzz_dbl_nc<P>(run);     // Vs = m * V
  store_xyzz<P>(V + t, run);
}
// The same level with one 4-lane TEAM per output entry: ~3.5x shorter dependent chain per addition.  Used for the
// window group that finishes last, where the latency of the chain is what the caller waits for; the single-thread
// version above has the better throughput (the upper groups' reductions run under other groups' accumulation).
template <class C>
__global__ void __launch_bounds__(128)
k_reduce_first_team(const XyzzMem<typename C::Fp>* __restrict__ buckets, int nslices, size_t slice_stride, size_t total_out,
                    int log_m, XyzzMem<typename C::Fp>* __restrict__ U, XyzzMem<typename C::Fp>* __restrict__ V) {
  using P = typename C::Fp;
  size_t t = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 2;
  if (t >= total_out) return;
  const Team tm;
  const XyzzMem<P>* b = buckets + (t << log_m);
  Xyzz<P> run = xyzz_inf<P>(), acc = xyzz_inf<P>();
  for (int i = (1 << log_m) - 1; i >= 0; i--) {
#pragma unroll 1
    for (int k = 0; k < nslices; k++) xyzz_add_tm<P>(tm, run, load_xyzz<P>(b + (size_t)k * slice_stride + i));
    xyzz_add_tm<P>(tm, acc, run);
  }
#pragma unroll 1
  for (int d = 0; d < log_m; d++) xyzz_dbl_tm<P>(tm, run);  // Vs = m * V
  if (tm.t == 0) { store_xyzz<P>(U + t, acc); store_xyzz<P>(V + t, run); }
}
// Next levels: groups of m entries (t = g*m + i):
//   U'[g] = sum_i U[t] + sum_i i*Vs[t],   Vs'[g] = m * sum_i Vs[t].
// One task (= one output entry) is worked on by 16 lanes = 4 teams of 4 lanes.  Three teams run the three
// chains of the level concurrently -- team 0: run += Vs_i, team 1: acc += run (one step behind, the run
// values are handed over through shared memory), team 2: usum += U_i -- so a level is m + 1 dependent
// additions deep instead of 3m, and the single inlined addition keeps operands in registers.  At the end team 0
// doubles its sum log2(m) times (the next level's Vs) while team 2 adds team 1's weighted sum to usum.
// `levels` > 1: the same CTA-resident tasks repeat the step on their own outputs (only when ONE block holds a whole
// segment's entries: the upper, tiny levels of the tree without a relaunch) -- see launch_reduce_next.
template <class C>
__global__ void __launch_bounds__(128)
k_reduce_next(const XyzzMem<typename C::Fp>* __restrict__ Uin, const XyzzMem<typename C::Fp>* __restrict__ Vin,
              size_t total_out, int log_m, XyzzMem<typename C::Fp>* __restrict__ Uout,
              XyzzMem<typename C::Fp>* __restrict__ Vout) {
  using P = typename C::Fp;
  __shared__ XyzzMem<P> hand[128 / 16][2];   // per task: double-buffered hand-over slot
  const size_t g = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 4;
  if (g >= total_out) return;
  const Team tm;
  const int lane = threadIdx.x & 31;
  const int q = (lane >> 2) & 3;                          // team within the task
  const unsigned gmask = 0xFFFFu << (lane & 16);          // the 16 lanes of this task
  XyzzMem<P>* slot = hand[(threadIdx.x >> 4)];
  const int m = 1 << log_m;
  const XyzzMem<P>* u = Uin + (g << log_m);
  const XyzzMem<P>* v = Vin + (g << log_m);
  Xyzz<P> X = xyzz_inf<P>();                              // this team's chain value
  for (int s = 0; s <= m; s++) {
    Xyzz<P> B = xyzz_inf<P>();
    if (s < m) {
      const int i = m - 1 - s;
      if (q == 0) B = load_xyzz<P>(v + i);
      else if (q == 1) { if (s >= 1) B = load_xyzz<P>(&slot[(s - 1) & 1]); }
      else if (q == 2) B = load_xyzz<P>(u + i);
    } else {
      // last step: team 1 hands its weighted sum to team 2, which adds it to usum; team 0 scales its plain sum by m
      if (q == 1 && tm.t == 0) store_xyzz<P>(&slot[0], X);
      __syncwarp(gmask);
      if (q == 2) B = load_xyzz<P>(&slot[0]);
    }
    if (s < m || q == 2) X = xyzz_add_team<P>(tm, X, B);  // the only addition site; B = infinity is a no-op
    if (s == m && q == 0) {
#pragma unroll 1
      for (int d = 0; d < log_m; d++) X = xyzz_dbl_team<P>(tm, X);
    }
    if (s < m && q == 0 && tm.t == 0) store_xyzz<P>(&slot[s & 1], X);
    __syncwarp(gmask);
  }
  if (tm.t == 0) {
    if (q == 2) store_xyzz<P>(Uout + g, X);
    if (q == 0) store_xyzz<P>(Vout + g, X);
  }
}

// ---- K5', the low-latency form of the bucket reduction (used where a caller waits for it: single MSMs) -------------
// The level recurrence above is m + 1 + log2(m) dependent group operations per level, and its upper levels are tiny:
// five launches of ~65 us each that nothing can fill.  Here the NB = 2^(c-1) buckets of a window are seen as a
// 2^hr x 2^hc matrix (bucket k = k1 * 2^hc + k0):
//     sum_k (k+1) B_k  =  2^hc * sum_k1 k1 * Row_k1  +  sum_k0 k0 * Col_k0  +  sum_k B_k
// with plain row and column sums (k_red_rowcol: every bucket is added twice, like the running sums do, but as short
// serial pieces plus a tree inside the block), and the two small weighted sums are taken bit by bit,
//     sum_j j * X_j = sum_b 2^b * T_b,   T_b = sum of the X_j whose index has bit b set      (k_red_bits: trees again),
// so that what is left is ONE Horner chain over the c-1 bit positions per window (k_tail_group_bits: cut into pieces that
// several teams evaluate at the same time, all windows of the group at once), then the ordinary Horner over the windows.
// Three launches and about 13 + 9 + (c-1)/4 + 3 dependent additions (c = 16: 29) instead of ~65; measured under ncu for two
// BLS12-381 windows 185 + 63 + 184 us against 143 + 309 + 80 (profiles/r2_notes.md section 10).  All additions are 4-lane
// team additions.

// Sum of the partial values of each group of `tpo` consecutive teams of the block (tpo a power of two, uniform over the
// block); the result is valid in the group's first team.  (Compacting the surviving teams into the first warps on every
// level -- so that whole warps drop out -- was measured and lost: 185 -> 213 us for k_red_rowcol on two BLS12-381 windows;
// the kernel is bound by the latency of its 13 dependent additions and the extra shared-memory round trip per level
// costs more than the idle teams' pipe time.)
template <class P>
ZK_D Xyzz<P> red_block_tree(const Team& tm, Xyzz<P> acc, int p, int tpo, XyzzMem<P>* sm) {
  const int tq = threadIdx.x >> 2;
  for (int s = 1; s < tpo; s <<= 1) {
    if ((p & (2 * s - 1)) == s && tm.t == 0) store_xyzz<P>(&sm[tq], acc);
    __syncthreads();
    if ((p & (2 * s - 1)) == 0) xyzz_add_tm<P>(tm, acc, load_xyzz<P>(&sm[tq + s]));
    __syncthreads();
  }
  return acc;
}
// blocks [0, row_blocks): row sums, the rest: column sums; blockIdx.y = window of the group.
// RC[window][0 .. NR) = rows, RC[window][NR .. NR + NC) = columns.
template <class C>
__global__ void __launch_bounds__(128)
k_red_rowcol(const XyzzMem<typename C::Fp>* __restrict__ buckets, int nslices, size_t slice_stride, RedPlan pl,
             XyzzMem<typename C::Fp>* __restrict__ RC) {
  using P = typename C::Fp;
  __shared__ XyzzMem<P> sm[32];
  const Team tm;
  const RedTask t = red_rowcol_task(pl, blockIdx.x, threadIdx.x >> 2);
  const size_t NB = (size_t)1 << (pl.hr + pl.hc), nrc = ((size_t)1 << pl.hr) + ((size_t)1 << pl.hc);
  const XyzzMem<P>* b = buckets + (size_t)blockIdx.y * NB + t.first;
  Xyzz<P> acc = xyzz_inf<P>();
  for (uint32_t i = t.part; t.valid && i < t.entries; i += t.tpo) {
#pragma unroll 1
    for (int k = 0; k < nslices; k++) xyzz_add_tm<P>(tm, acc, load_xyzz<P>(b + (size_t)k * slice_stride + (size_t)i * t.stride));
  }
  acc = red_block_tree<P>(tm, acc, t.part, t.tpo, sm);
  if (t.valid && t.part == 0 && tm.t == 0)
    store_xyzz<P>(RC + (size_t)blockIdx.y * nrc + (t.rows ? t.out : ((size_t)1 << pl.hr) + t.out), acc);
}
// blockIdx.x = j: bit j of the column index (j < hc), bit j - hc of the row index (j < hc + hr), or the sum of all rows
// (j = hc + hr); blockIdx.y = window.  T[window][j].
template <class C>
__global__ void __launch_bounds__(128)
k_red_bits(const XyzzMem<typename C::Fp>* __restrict__ RC, RedPlan pl, XyzzMem<typename C::Fp>* __restrict__ T) {
  using P = typename C::Fp;
  __shared__ XyzzMem<P> sm[32];
  const Team tm;
  const int tq = threadIdx.x >> 2;
  const size_t nrc = ((size_t)1 << pl.hr) + ((size_t)1 << pl.hc);
  const int j = blockIdx.x;
  const RedBits bt = red_bits_task(pl, j);
  const XyzzMem<P>* src = RC + (size_t)blockIdx.y * nrc + bt.base;
  Xyzz<P> acc = xyzz_inf<P>();
  for (uint32_t e = tq; e < bt.entries; e += 32) xyzz_add_tm<P>(tm, acc, load_xyzz<P>(src + red_bit_member(e, bt.bit)));
  acc = red_block_tree<P>(tm, acc, tq, 32, sm);
  if (tq == 0 && tm.t == 0) store_xyzz<P>(T + (size_t)blockIdx.y * (pl.hr + pl.hc + 1) + j, acc);
}
// One block, NCH teams per window of the group.  The window sum  sum_b 2^b T_b + T_all  is a Horner chain over the bit
// positions; NCH teams take NCH contiguous pieces of it at the same time (piece q covers bits [lo_q, hi_q), its value is
// sum_{b in piece} 2^(b - lo_q) T_b), then the window's first team joins the pieces top down (hi_q - lo_q doublings and
// one addition each) -- the chain is (c-1)/NCH + NCH - 1 additions deep instead of c - 1.  Then team 0 of the block:
// Horner over the windows (top first) and `extra` more doublings, as k_tail_group.
template <class C, int NCH>
__global__ void k_tail_group_bits(const XyzzMem<typename C::Fp>* __restrict__ T, int Wg, int c, int extra,
                                  XyzzMem<typename C::Fp>* __restrict__ out) {
  using P = typename C::Fp;
  __shared__ XyzzMem<P> Rw[32];                   // piece values, then (slot w * NCH) the window sums
  const Team tm;
  const int tq = threadIdx.x >> 2, w = tq / NCH, q = tq % NCH;
  const int nb = c - 1;                           // bit positions 0 .. nb-1
  int lo, hi;
  red_piece(nb, NCH, q, lo, hi);
  const XyzzMem<P>* t = T + (size_t)w * c;        // c - 1 bit sums and the plain sum
  if (w < Wg) {                                   // whole teams take the branch
    Xyzz<P> acc = xyzz_inf<P>();
#pragma unroll 1
    for (int b = hi - 1; b >= lo; b--) {
      acc = xyzz_dbl_team<P>(tm, acc);            // no-op while acc is infinity
      xyzz_add_tm<P>(tm, acc, load_xyzz<P>(t + b));
    }
    if (q == 0) xyzz_add_tm<P>(tm, acc, load_xyzz<P>(t + (c - 1)));
    if (NCH > 1 && tm.t == 0) store_xyzz<P>(&Rw[tq], acc);
    if (NCH > 1) __syncwarp(0xFFFFFFFFu >> (32 - 4 * NCH) << (4 * (tq - q) & 31));   // the window's teams share a warp
    if (q == 0) {
      if (NCH > 1) {
        acc = load_xyzz<P>(&Rw[tq + NCH - 1]);
#pragma unroll 1
        for (int r = NCH - 2; r >= 0; r--) {
          int rlo, rhi;
          red_piece(nb, NCH, r, rlo, rhi);
          const int sh = rhi - rlo;                 // width of piece r: what is above it moves up by that much
#pragma unroll 1
          for (int d = 0; d < sh; d++) acc = xyzz_dbl_team<P>(tm, acc);
          xyzz_add_tm<P>(tm, acc, load_xyzz<P>(&Rw[tq + r]));
        }
      }
      __syncwarp(tm.mask);
      if (tm.t == 0) store_xyzz<P>(&Rw[tq], acc);
    }
  }
  __syncthreads();
  if (threadIdx.x >= 4) return;
  Xyzz<P> acc = xyzz_inf<P>();
  for (int v = Wg - 1; v >= 0; v--) {
    const int nd = v == Wg - 1 ? 0 : c;
#pragma unroll 1
    for (int d = 0; d < nd; d++) acc = xyzz_dbl_team<P>(tm, acc);
    xyzz_add_tm<P>(tm, acc, load_xyzz<P>(&Rw[v * NCH]));
  }
#pragma unroll 1
  for (int d = 0; d < extra; d++) acc = xyzz_dbl_team<P>(tm, acc);
  if (tm.t == 0) store_xyzz<P>(out, acc);
}
// buckets of `ns` consecutive windows -> the group's share of the result (XYZZ) in *out.
// Workspaces: RC ns * (2^hr + 2^hc) records, T ns * c records.
template <class C>
int launch_reduce_2d(cudaStream_t s, const XyzzMem<typename C::Fp>* buckets, int nslices, size_t slice_stride, int ns, int c, int extra,
                     XyzzMem<typename C::Fp>* RC, XyzzMem<typename C::Fp>* T, XyzzMem<typename C::Fp>* out) {
  const RedPlan pl = red_plan(c);
  k_red_rowcol<C><<<dim3(pl.row_blocks + pl.col_blocks, ns), 128, 0, s>>>(buckets, nslices, slice_stride, pl, RC);
  k_red_bits<C><<<dim3(c, ns), 128, 0, s>>>(RC, pl, T);
  // pieces per window: as many as fit into one block of 32 teams (a window's teams must share a warp: 8 teams)
  if (ns <= 8) k_tail_group_bits<C, 4><<<1, ((16 * ns + 31) / 32) * 32, 0, s>>>(T, ns, c, extra, out);
  else if (ns <= 16) k_tail_group_bits<C, 2><<<1, ((8 * ns + 31) / 32) * 32, 0, s>>>(T, ns, c, extra, out);
  else k_tail_group_bits<C, 1><<<1, ((4 * ns + 31) / 32) * 32, 0, s>>>(T, ns, c, extra, out);
  return 3;
}

// ---- K6 + K7: window combination (Horner) and output conversion ------------------------------------------

template <class P>
ZK_D void write_fe(uint32_t* dst, const Fe<P>& a) {
#pragma unroll
  for (int i = 0; i < P::L; i++) dst[i] = a.l[i];
}
template <class P>
ZK_D Fe<P> read_fe(const uint32_t* src) {
  Fe<P> a;
#pragma unroll
  for (int i = 0; i < P::L; i++) a.l[i] = src[i];
  return a;
}
template <class P>
ZK_D void write_result(uint32_t* o, const Xyzz<P>& acc, int mode) {
  constexpr int L = P::L;
  if (mode == OUT_AFFINE) {
    Affine<P> a;
    if (xyzz_to_affine<P>(acc, a)) { write_fe<P>(o, a.x); write_fe<P>(o + L, a.y); }
    else { for (int i = 0; i < 2 * L; i++) o[i] = 0xffffffffu; }
  } else if (mode == OUT_XYZZ) {
    write_fe<P>(o, acc.X); write_fe<P>(o + L, acc.Y); write_fe<P>(o + 2 * L, acc.ZZ); write_fe<P>(o + 3 * L, acc.ZZZ);
  } else {
    Fe<P> X, Y, Z;
    if (mode == OUT_PROJ) xyzz_to_proj<P>(acc, X, Y, Z); else xyzz_to_jac<P>(acc, X, Y, Z);
    write_fe<P>(o, X); write_fe<P>(o + L, Y); write_fe<P>(o + 2 * L, Z);
  }
}
// one 4-lane team per MSM of the batch; Rw[msm*W + w] = window sums; top-down Horner: acc = 2^c * acc + R_w
// (c inlined team doublings per window); then output conversion (record stride 4L words).
template <class C>
__global__ void __launch_bounds__(32)
k_tail(const XyzzMem<typename C::Fp>* __restrict__ Rw, int nmsm, int W, int c, int mode, uint32_t* __restrict__ out) {
  using P = typename C::Fp;
  int m = (blockIdx.x * blockDim.x + threadIdx.x) >> 2;
  if (m >= nmsm) return;
  Team tm;
  Xyzz<P> acc = xyzz_inf<P>();
  for (int w = W - 1; w >= 0; w--) {
#pragma unroll 1
    for (int d = 0; d < c; d++) acc = xyzz_dbl_team<P>(tm, acc);   // inlined once; no-op while acc is infinity
    xyzz_add_tm<P>(tm, acc, load_xyzz<P>(Rw + (size_t)m * W + w));
  }
  if (tm.t == 0) write_result<P>(out + (size_t)m * (4 * P::L), acc, mode);
}

// Horner over the Wg window sums of ONE group of consecutive windows (top window first), then `extra` more
// doublings = c * (index of the group's lowest window): the group's share of the result, stored as XYZZ.
// Lets the combination of the upper windows run while the lower windows are still being accumulated.
template <class C>
__global__ void k_tail_group(const XyzzMem<typename C::Fp>* __restrict__ Rw, int Wg, int c, int extra,
                             XyzzMem<typename C::Fp>* __restrict__ out) {
  using P = typename C::Fp;
  if (blockIdx.x != 0 || threadIdx.x >= 4) return;
  Team tm;
  Xyzz<P> acc = xyzz_inf<P>();
  for (int w = Wg - 1; w >= 0; w--) {
    const int nd = w == Wg - 1 ? 0 : c;
#pragma unroll 1
    for (int d = 0; d < nd; d++) acc = xyzz_dbl_team<P>(tm, acc);
    xyzz_add_tm<P>(tm, acc, load_xyzz<P>(Rw + w));
  }
#pragma unroll 1
  for (int d = 0; d < extra; d++) acc = xyzz_dbl_team<P>(tm, acc);
  if (tm.t == 0) store_xyzz<P>(out, acc);
}

// sum of k group elements given in one of the reference's representations (multi-GPU combine, K8)
//   in_mode: OUT_PROJ / OUT_JAC / OUT_XYZZ (records of 3L / 3L / 4L words)
template <class C>
__global__ void k_sum_points(const uint32_t* __restrict__ in, int k, int in_mode, int out_mode, uint32_t* __restrict__ out) {
  using P = typename C::Fp;
  constexpr int L = P::L;
  if (blockIdx.x != 0 || threadIdx.x >= 4) return;
  Team tm;
  Xyzz<P> acc = xyzz_inf<P>();
  for (int i = 0; i < k; i++) {
    Xyzz<P> p;
    if (in_mode == OUT_XYZZ) {
      const uint32_t* s = in + (size_t)i * 4 * L;
      p.X = read_fe<P>(s); p.Y = read_fe<P>(s + L); p.ZZ = read_fe<P>(s + 2 * L); p.ZZZ = read_fe<P>(s + 3 * L);
    } else {
      const uint32_t* s = in + (size_t)i * 3 * L;
      Fe<P> X = read_fe<P>(s), Y = read_fe<P>(s + L), Z = read_fe<P>(s + 2 * L);
      p = (in_mode == OUT_PROJ) ? xyzz_from_proj<P>(X, Y, Z) : xyzz_from_jac<P>(X, Y, Z);
    }
    xyzz_add_tm<P>(tm, acc, p);
  }
  if (tm.t == 0) write_result<P>(out, acc, out_mode);
}


// Workload synthesis (not on the MSM path): out[i] = P0 + (start + i) * D as canonical affine points.
// One thread per point: double-and-add of the index, one mixed add, one inversion.
template <class C>
__global__ void __launch_bounds__(128)
k_gen_chain(const uint32_t* __restrict__ p0d, unsigned long long start, size_t n, uint32_t* __restrict__ out) {
  using P = typename C::Fp;
  constexpr int L = P::L;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Affine<P> p0, d;
  p0.x = read_fe<P>(p0d); p0.y = read_fe<P>(p0d + L); d.x = read_fe<P>(p0d + 2 * L); d.y = read_fe<P>(p0d + 3 * L);
  unsigned long long k = start + i;
  Xyzz<P> acc = xyzz_inf<P>();
  for (int b = 63; b >= 0; b--) {
    xyzz_dbl_nc<P>(acc);
    if ((k >> b) & 1ull) { Xyzz<P> t = xyzz_from_affine<P>(d); xyzz_add_nc<P>(acc, t); }
  }
  { Xyzz<P> t = xyzz_from_affine<P>(p0); xyzz_add_nc<P>(acc, t); }
  uint32_t* o = out + i * (2 * L);
  Affine<P> a;
  if (xyzz_to_affine<P>(acc, a)) { write_fe<P>(o, a.x); write_fe<P>(o + L, a.y); }
  else { for (int j = 0; j < 2 * L; j++) o[j] = 0xffffffffu; }
}

// ---- next row of the scope table: batch conversions -------------------------------
// <curve>_G1_{proj,jac}_batch_to_affine: the reference does N separate inversions
// (lib/cbits/curves/g1/proj/bn128_G1_proj.c:158-166 -> :132-144; jac: bn128_G1_jac.c:147-155 -> :120-136).
// Here every thread converts BATCH consecutive points with ONE inversion (Montgomery's trick); Z = 0 gives
// the all-0xFF record exactly like the reference.  Output is canonical, hence bit-identical.
constexpr int TOAFF_BATCH = 4;
template <class C, bool JAC>
__global__ void __launch_bounds__(128)
k_batch_to_affine(const uint32_t* __restrict__ src, size_t n, uint32_t* __restrict__ dst) {
  using P = typename C::Fp;
  constexpr int L = P::L;
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t first = t * TOAFF_BATCH;
  if (first >= n) return;
  int cnt = (int)((n - first) < (size_t)TOAFF_BATCH ? (n - first) : (size_t)TOAFF_BATCH);
  Fe<P> z[TOAFF_BATCH], pre[TOAFF_BATCH];
  Fe<P> acc = fe_one<P>();
  for (int k = 0; k < cnt; k++) {
    z[k] = read_fe<P>(src + (first + k) * 3 * L + 2 * L);
    pre[k] = acc;
    if (!fe_is_zero<P>(z[k])) acc = fe_mul_call<P>(acc, z[k]);
  }
  Fe<P> inv = fe_inv<P>(acc);   // acc is a product of non-zero elements (or one)
  for (int k = cnt - 1; k >= 0; k--) {
    uint32_t* o = dst + (first + k) * 2 * L;
    if (fe_is_zero<P>(z[k])) {
      for (int i = 0; i < 2 * L; i++) o[i] = 0xffffffffu;
      continue;
    }
    Fe<P> zi = fe_mul_call<P>(inv, pre[k]);
    inv = fe_mul_call<P>(inv, z[k]);
    const uint32_t* s = src + (first + k) * 3 * L;
    Fe<P> X = read_fe<P>(s), Y = read_fe<P>(s + L);
    if (JAC) {
      Fe<P> zi2 = fe_mul_call<P>(zi, zi);
      write_fe<P>(o, fe_mul_call<P>(X, zi2));
      write_fe<P>(o + L, fe_mul_call<P>(Y, fe_mul_call<P>(zi2, zi)));
    } else {
      write_fe<P>(o, fe_mul_call<P>(X, zi));
      write_fe<P>(o + L, fe_mul_call<P>(Y, zi));
    }
  }
}
// The same conversion for LARGE arrays with ONE field inversion per call: the Z coordinates (1 where Z = 0) go through the
// batch inversion tree of the affine pre-reduction (kernels_aff.cuh: batch_invert, 3 multiplications per element), then
// every point is scaled by its own 1/Z.  2 + 3 (proj) or 4 + 3 (jac) multiplications per point instead of a quarter of a
// full inversion.  Same canonical output.
template <class C>
__global__ void __launch_bounds__(256)
k_convert_z(const uint32_t* __restrict__ src, size_t n, uint32_t* __restrict__ z_out) {
  using P = typename C::Fp;
  constexpr int L = P::L;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fe<P> z = read_fe<P>(src + i * 3 * L + 2 * L);
  if (fe_is_zero<P>(z)) z = fe_one<P>();
  write_fe<P>(z_out + i * L, z);
}
template <class C, bool JAC>
__global__ void __launch_bounds__(128)
k_convert_apply(const uint32_t* __restrict__ src, const uint32_t* __restrict__ zinv, size_t n, uint32_t* __restrict__ dst) {
  using P = typename C::Fp;
  constexpr int L = P::L;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t* s = src + i * 3 * L;
  uint32_t* o = dst + i * 2 * L;
  if (fe_is_zero<P>(read_fe<P>(s + 2 * L))) {
    for (int k = 0; k < 2 * L; k++) o[k] = 0xffffffffu;
    return;
  }
  Fe<P> zi = read_fe<P>(zinv + i * L), X = read_fe<P>(s), Y = read_fe<P>(s + L);
  if (JAC) {
    Fe<P> zi2 = fe_mul_call<P>(zi, zi);
    write_fe<P>(o, fe_mul_call<P>(X, zi2));
    write_fe<P>(o + L, fe_mul_call<P>(Y, fe_mul_call<P>(zi2, zi)));
  } else {
    write_fe<P>(o, fe_mul_call<P>(X, zi));
    write_fe<P>(o + L, fe_mul_call<P>(Y, zi));
  }
}
// <curve>_G1_{proj,jac}_batch_from_affine (bn128_G1_proj.c:147-155 -> :120-128; bn128_G1_jac.c:138-145 -> :105-116)
template <class C, bool JAC>
__global__ void __launch_bounds__(256)
k_batch_from_affine(const uint32_t* __restrict__ src, size_t n, uint32_t* __restrict__ dst) {
  using P = typename C::Fp;
  constexpr int L = P::L;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Affine<P> a;
  const uint32_t* s = src + i * 2 * L;
  a.x = read_fe<P>(s); a.y = read_fe<P>(s + L);
  uint32_t* o = dst + i * 3 * L;
  if (affine_is_inf<P>(a)) {
    write_fe<P>(o, JAC ? fe_one<P>() : fe_zero<P>());
    write_fe<P>(o + L, fe_one<P>());
    write_fe<P>(o + 2 * L, fe_zero<P>());
  } else {
    write_fe<P>(o, a.x); write_fe<P>(o + L, a.y); write_fe<P>(o + 2 * L, fe_one<P>());
  }
}
template <class C>
void launch_batch_to_affine(cudaStream_t s, const uint32_t* src, size_t n, uint32_t* dst, int jac) {
  size_t threads = (n + TOAFF_BATCH - 1) / TOAFF_BATCH;
  if (jac) k_batch_to_affine<C, true><<<(unsigned)((threads + 127) / 128), 128, 0, s>>>(src, n, dst);
  else k_batch_to_affine<C, false><<<(unsigned)((threads + 127) / 128), 128, 0, s>>>(src, n, dst);
}
template <class C>
void launch_convert_z(cudaStream_t s, const uint32_t* src, size_t n, uint32_t* z_out) {
  k_convert_z<C><<<(unsigned)((n + 255) / 256), 256, 0, s>>>(src, n, z_out);
}
template <class C>
void launch_convert_apply(cudaStream_t s, const uint32_t* src, const uint32_t* zinv, size_t n, uint32_t* dst, int jac) {
  if (jac) k_convert_apply<C, true><<<(unsigned)((n + 127) / 128), 128, 0, s>>>(src, zinv, n, dst);
  else k_convert_apply<C, false><<<(unsigned)((n + 127) / 128), 128, 0, s>>>(src, zinv, n, dst);
}
template <class C>
void launch_batch_from_affine(cudaStream_t s, const uint32_t* src, size_t n, uint32_t* dst, int jac) {
  if (jac) k_batch_from_affine<C, true><<<(unsigned)((n + 255) / 256), 256, 0, s>>>(src, n, dst);
  else k_batch_from_affine<C, false><<<(unsigned)((n + 255) / 256), 256, 0, s>>>(src, n, dst);
}

template <class C>
void launch_gen_chain(cudaStream_t s, const uint32_t* p0d, unsigned long long start, size_t n, uint32_t* out) {
  k_gen_chain<C><<<(unsigned)((n + 127) / 128), 128, 0, s>>>(p0d, start, n, out);
}

template <class C>
void launch_reduce_first(cudaStream_t s, const XyzzMem<typename C::Fp>* buckets, int nslices, size_t slice_stride, size_t total_out,
                         int log_m, XyzzMem<typename C::Fp>* U, XyzzMem<typename C::Fp>* V, int team) {
  // 64-thread blocks: these launches run on high-priority side streams WHILE the accumulation kernels of other window
  // groups own the SMs (3 CTAs of ~19.5 K registers each); a block is scheduled as soon as ONE of those CTAs retires only
  // if its registers fit into that hole -- 128 threads x 255 registers never did, so the "overlapped" reductions used to
  // wait until the accumulation queue ran dry.
  if (team) k_reduce_first_team<C><<<(unsigned)((total_out * 4 + 63) / 64), 64, 0, s>>>(buckets, nslices, slice_stride, total_out, log_m, U, V);
  else k_reduce_first<C><<<(unsigned)((total_out + 63) / 64), 64, 0, s>>>(buckets, nslices, slice_stride, total_out, log_m, U, V);
}
template <class C>
void launch_reduce_next(cudaStream_t s, const XyzzMem<typename C::Fp>* Uin, const XyzzMem<typename C::Fp>* Vin, size_t total_out,
                        int log_m, XyzzMem<typename C::Fp>* Uout, XyzzMem<typename C::Fp>* Vout) {
  k_reduce_next<C><<<(unsigned)((total_out * 16 + 63) / 64), 64, 0, s>>>(Uin, Vin, total_out, log_m, Uout, Vout);
}
template <class C>
void launch_tail(cudaStream_t s, const XyzzMem<typename C::Fp>* Rw, int nmsm, int W, int c, int mode, uint32_t* out) {
  k_tail<C><<<(nmsm * 4 + 31) / 32, 32, 0, s>>>(Rw, nmsm, W, c, mode, out);
}
template <class C>
void launch_tail_group(cudaStream_t s, const XyzzMem<typename C::Fp>* Rw, int Wg, int c, int extra, XyzzMem<typename C::Fp>* out) {
  k_tail_group<C><<<1, 32, 0, s>>>(Rw, Wg, c, extra, out);
}
template <class C>
void launch_sum_points(cudaStream_t s, const uint32_t* in, int k, int in_mode, int out_mode, uint32_t* out) {
  k_sum_points<C><<<1, 32, 0, s>>>(in, k, in_mode, out_mode, out);
}

#define ZK_INSTANTIATE_RED(C)                                                                                             \
  template void launch_reduce_first<C>(cudaStream_t, const XyzzMem<C::Fp>*, int, size_t, size_t, int, XyzzMem<C::Fp>*,     \
                                       XyzzMem<C::Fp>*, int);                                                             \
  template void launch_reduce_next<C>(cudaStream_t, const XyzzMem<C::Fp>*, const XyzzMem<C::Fp>*, size_t, int,             \
                                      XyzzMem<C::Fp>*, XyzzMem<C::Fp>*);                                                   \
  template void launch_tail<C>(cudaStream_t, const XyzzMem<C::Fp>*, int, int, int, int, uint32_t*);                        \
  template void launch_sum_points<C>(cudaStream_t, const uint32_t*, int, int, int, uint32_t*);                             \
  template void launch_tail_group<C>(cudaStream_t, const XyzzMem<C::Fp>*, int, int, int, XyzzMem<C::Fp>*);                 \
  template int launch_reduce_2d<C>(cudaStream_t, const XyzzMem<C::Fp>*, int, size_t, int, int, int, XyzzMem<C::Fp>*,       \
                                   XyzzMem<C::Fp>*, XyzzMem<C::Fp>*);                                                      \
  template void launch_gen_chain<C>(cudaStream_t, const uint32_t*, unsigned long long, size_t, uint32_t*);               \
  template void launch_batch_to_affine<C>(cudaStream_t, const uint32_t*, size_t, uint32_t*, int);                         \
  template void launch_convert_z<C>(cudaStream_t, const uint32_t*, size_t, uint32_t*);                                     \
  template void launch_convert_apply<C>(cudaStream_t, const uint32_t*, const uint32_t*, size_t, uint32_t*, int);           \
  template void launch_batch_from_affine<C>(cudaStream_t, const uint32_t*, size_t, uint32_t*, int);

}  // namespace zk
