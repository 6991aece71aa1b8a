// Segmented LSD radix sort of (bucket key, point index) pairs, 8 bits per pass.
//
// This is the explicit form of the reference's implicit "scatter every point into SUMS[e]"
//   lib/cbits/curves/g1/proj/bn128_G1_proj.c:549-561
// One segment = one Pippenger window (of one MSM of a batch); segments are sorted independently,
// all in the same launches (blockIdx.y = segment).
//
// Per pass, three kernels (deterministic, stable):
//   k_sort_hist     tile digit histogram          -> cnt[seg][digit][tile]
//   k_sort_rowscan  exclusive scan along tiles    -> cnt in place, rowsum[seg][digit]
//   k_sort_scatter  128-bit loads -> shared memory, stable in-tile rank (warp match), staged in shared
//                   memory in sorted order, copied out coalesced per digit run
// HBM traffic per pass and pair: 4 B (hist) + 8 B read + 8 B written.
#include "sort.cuh"

namespace zk {

__global__ void __launch_bounds__(SORT_THREADS)
k_sort_hist(const uint32_t* __restrict__ keys, size_t n, int shift, uint32_t* __restrict__ cnt, int tiles) {
  __shared__ uint32_t h[SORT_RADIX];
  const int seg = blockIdx.y, tile = blockIdx.x;
  if (threadIdx.x < SORT_RADIX) h[threadIdx.x] = 0;
  __syncthreads();
  const uint32_t* k = keys + (size_t)seg * n;
  const size_t base = (size_t)tile * SORT_TILE;
  const uint32_t count = (uint32_t)((n - base) < (size_t)SORT_TILE ? (n - base) : (size_t)SORT_TILE);
  if ((((uintptr_t)(k + base)) & 15u) == 0) {
    for (uint32_t i = threadIdx.x * 4; i < SORT_TILE; i += SORT_THREADS * 4) {
      if (i + 3 < count) {
        uint4 v = *reinterpret_cast<const uint4*>(k + base + i);
        atomicAdd(&h[(v.x >> shift) & 0xffu], 1u);
        atomicAdd(&h[(v.y >> shift) & 0xffu], 1u);
        atomicAdd(&h[(v.z >> shift) & 0xffu], 1u);
        atomicAdd(&h[(v.w >> shift) & 0xffu], 1u);
      } else {
        for (uint32_t q = i; q < i + 4 && q < count; q++) atomicAdd(&h[(k[base + q] >> shift) & 0xffu], 1u);
      }
    }
  } else {
    for (uint32_t i = threadIdx.x; i < count; i += SORT_THREADS) atomicAdd(&h[(k[base + i] >> shift) & 0xffu], 1u);
  }
  __syncthreads();
  if (threadIdx.x < SORT_RADIX) cnt[((size_t)seg * SORT_RADIX + threadIdx.x) * tiles + tile] = h[threadIdx.x];
}

// grid (256 digits, segments); exclusive scan of one row of `tiles` counters, total -> rowsum
__global__ void __launch_bounds__(256)
k_sort_rowscan(uint32_t* __restrict__ cnt, uint32_t* __restrict__ rowsum, int tiles) {
  __shared__ uint32_t wsum[8];
  __shared__ uint32_t carry_s;
  const int digit = blockIdx.x, seg = blockIdx.y;
  uint32_t* row = cnt + ((size_t)seg * SORT_RADIX + digit) * tiles;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < tiles; base += 256) {
    int i = base + threadIdx.x;
    uint32_t v = (i < tiles) ? row[i] : 0u;
    uint32_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) wsum[warp] = x;
    __syncthreads();
    uint32_t woff = 0;
#pragma unroll
    for (int w = 0; w < 8; w++) woff += (w < warp) ? wsum[w] : 0u;
    uint32_t carry = carry_s;
    if (i < tiles) row[i] = carry + woff + x - v;
    __syncthreads();
    if (threadIdx.x == 255) carry_s = carry + woff + x;
    __syncthreads();
  }
  if (threadIdx.x == 0) rowsum[(size_t)seg * SORT_RADIX + digit] = carry_s;
}

// Scatter kernel with shared-memory staging on both sides:
//   1. the tile's 4096 pairs are read with 128-bit loads into shared memory;
//   2. every warp ranks its items in index order (stable): ballots group the lanes with equal digits,
//      the group leader bumps the warp's digit counter;
//   3. per-digit scan over the warps + scan over the digits give each item its slot in the tile's
//      sorted order; the items are written to that slot in shared memory;
//   4. the staged tile is copied out in slot order, so that consecutive threads write consecutive
//      addresses inside every digit's run (average run: 16 pairs).
__global__ void __launch_bounds__(SORT_THREADS)
k_sort_scatter(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
               uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out, size_t n, int shift,
               const uint32_t* __restrict__ cnt, const uint32_t* __restrict__ rowsum, int tiles) {
  __shared__ __align__(16) uint32_t s_key[SORT_TILE];
  __shared__ __align__(16) uint32_t s_val[SORT_TILE];
  __shared__ uint16_t wcnt[SORT_THREADS / 32][SORT_RADIX + 2];   // per-warp digit counters (<= 4096 each)
  __shared__ uint32_t dstart[SORT_RADIX];   // first slot of digit d in the staged tile
  __shared__ uint32_t gbase[SORT_RADIX];    // global position of that slot
  __shared__ uint32_t scan_tmp[SORT_RADIX];
  const int seg = blockIdx.y, tile = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t* kin = keys_in + (size_t)seg * n;
  const uint32_t* vin = vals_in + (size_t)seg * n;
  const size_t tbase = (size_t)tile * SORT_TILE;
  const uint32_t count = (uint32_t)((n - tbase) < (size_t)SORT_TILE ? (n - tbase) : (size_t)SORT_TILE);

  for (int i = tid; i < (SORT_THREADS / 32) * (SORT_RADIX + 2); i += SORT_THREADS) (&wcnt[0][0])[i] = 0;
  // ---- 1. load ----
  const bool aligned = (((uintptr_t)(kin + tbase) | (uintptr_t)(vin + tbase)) & 15u) == 0;
  if (aligned) {
    for (uint32_t i = tid * 4; i < SORT_TILE; i += SORT_THREADS * 4) {
      if (i + 3 < count) {
        *reinterpret_cast<uint4*>(s_key + i) = *reinterpret_cast<const uint4*>(kin + tbase + i);
        *reinterpret_cast<uint4*>(s_val + i) = *reinterpret_cast<const uint4*>(vin + tbase + i);
      } else {
        for (uint32_t q = i; q < i + 4 && q < count; q++) { s_key[q] = kin[tbase + q]; s_val[q] = vin[tbase + q]; }
      }
    }
  } else {
    for (uint32_t i = tid; i < count; i += SORT_THREADS) { s_key[i] = kin[tbase + i]; s_val[i] = vin[tbase + i]; }
  }
  // global base of every digit of this segment: exclusive scan of the 256 row totals
  const bool dthread = tid < SORT_RADIX;   // the threads that own one digit each
  uint32_t mine = dthread ? rowsum[(size_t)seg * SORT_RADIX + tid] : 0u;
  if (dthread) scan_tmp[tid] = mine;
  __syncthreads();
  for (int o = 1; o < SORT_RADIX; o <<= 1) {
    uint32_t y = (dthread && tid >= o) ? scan_tmp[tid - o] : 0u;
    __syncthreads();
    if (dthread) scan_tmp[tid] += y;
    __syncthreads();
  }
  if (dthread) gbase[tid] = scan_tmp[tid] - mine + cnt[((size_t)seg * SORT_RADIX + tid) * tiles + tile];

  // ---- 2. rank (items of warp w in striped order: index = w*512 + i*32 + lane) ----
  uint32_t key[SORT_ITEMS], val[SORT_ITEMS], off[SORT_ITEMS];
  const uint32_t wbase = warp * (32 * SORT_ITEMS);
#pragma unroll
  for (int i = 0; i < SORT_ITEMS; i++) {
    uint32_t idx = wbase + i * 32 + lane;
    bool ok = idx < count;
    key[i] = ok ? s_key[idx] : 0u;
    val[i] = ok ? s_val[idx] : 0u;
  }
  // all peer masks first (independent of each other), then the serial counter updates
  uint32_t peers[SORT_ITEMS];
#pragma unroll
  for (int i = 0; i < SORT_ITEMS; i++) {
    uint32_t idx = wbase + i * 32 + lane;
    uint32_t d = (idx < count) ? ((key[i] >> shift) & 0xffu) : (uint32_t)SORT_RADIX;  // tail lanes share a dummy bin
    // lanes with the same 9-bit value, from 9 ballots (MATCH.ANY costs one round per DISTINCT value in the
    // warp, ~30 here; the ballot form is a fixed 9 votes + 9 logic ops)
    uint32_t m = 0xffffffffu;
#pragma unroll
    for (int b = 0; b < 9; b++) {
      bool bit = (d >> b) & 1u;
      uint32_t bal = __ballot_sync(0xffffffffu, bit);
      m &= bit ? bal : ~bal;
    }
    peers[i] = m;
  }
#pragma unroll
  for (int i = 0; i < SORT_ITEMS; i++) {
    uint32_t idx = wbase + i * 32 + lane;
    uint32_t d = (idx < count) ? ((key[i] >> shift) & 0xffu) : (uint32_t)SORT_RADIX;
    int leader = __ffs(peers[i]) - 1;
    uint32_t before = __popc(peers[i] & ((1u << lane) - 1u));
    uint32_t old = 0;
    if (lane == leader) {
      old = wcnt[warp][d];
      wcnt[warp][d] = (uint16_t)(old + __popc(peers[i]));
    }
    old = __shfl_sync(0xffffffffu, old, leader);
    off[i] = old + before;
    __syncwarp();
  }
  __syncthreads();
  // ---- 3. slots ----
  uint32_t tot = 0;
  if (dthread) {
#pragma unroll
    for (int w = 0; w < SORT_THREADS / 32; w++) {
      uint32_t t = wcnt[w][tid];
      wcnt[w][tid] = (uint16_t)tot;
      tot += t;
    }
    scan_tmp[tid] = tot;
  }
  __syncthreads();
  for (int o = 1; o < SORT_RADIX; o <<= 1) {
    uint32_t y = (dthread && tid >= o) ? scan_tmp[tid - o] : 0u;
    __syncthreads();
    if (dthread) scan_tmp[tid] += y;
    __syncthreads();
  }
  if (dthread) dstart[tid] = scan_tmp[tid] - tot;
  __syncthreads();
#pragma unroll
  for (int i = 0; i < SORT_ITEMS; i++) {
    uint32_t idx = wbase + i * 32 + lane;
    if (idx < count) {
      uint32_t d = (key[i] >> shift) & 0xffu;
      uint32_t slot = dstart[d] + wcnt[warp][d] + off[i];
      s_key[slot] = key[i];
      s_val[slot] = val[i];
    }
  }
  __syncthreads();
  // ---- 4. copy out in slot order ----
  uint32_t* kout = keys_out + (size_t)seg * n;
  uint32_t* vout = vals_out + (size_t)seg * n;
  for (uint32_t i = tid; i < count; i += SORT_THREADS) {
    uint32_t k = s_key[i];
    uint32_t d = (k >> shift) & 0xffu;
    size_t pos = (size_t)gbase[d] + (i - dstart[d]);
    kout[pos] = k;
    vout[pos] = s_val[i];
  }
}

void sort_pass(cudaStream_t s, const uint32_t* keys_in, const uint32_t* vals_in, uint32_t* keys_out, uint32_t* vals_out,
               size_t n, int nseg, int shift, uint32_t* cnt, uint32_t* rowsum, int tiles) {
  dim3 g(tiles, nseg);
  k_sort_hist<<<g, SORT_THREADS, 0, s>>>(keys_in, n, shift, cnt, tiles);
  k_sort_rowscan<<<dim3(SORT_RADIX, nseg), 256, 0, s>>>(cnt, rowsum, tiles);
  k_sort_scatter<<<g, SORT_THREADS, 0, s>>>(keys_in, vals_in, keys_out, vals_out, n, shift, cnt, rowsum, tiles);
}

}  // namespace zk
