// Segmented LSD radix sort of packed (bucket key, point index) pairs, 8 bits per pass: see sort.cuh.
#include "sort.cuh"

namespace zk {

// tile digit histogram of the packed pairs -> cnt[seg][digit][tile]  (counting the next pass's digits from inside the
// scatter kernel with red.global was measured and lost: 0.32 -> 0.59 .. 1.48 ms per pass at 2^22 x 14 pairs, the upper
// digits have few distinct values and the atomics of a whole tile pile up on a handful of addresses)
__global__ void __launch_bounds__(SORT_THREADS)
k_sort_hist(const uint2* __restrict__ in, size_t n, int shift, uint32_t* __restrict__ cnt, int tiles) {
  __shared__ uint32_t h[SORT_RADIX];
  const int seg = blockIdx.y, tile = blockIdx.x;
  if (threadIdx.x < SORT_RADIX) h[threadIdx.x] = 0;
  __syncthreads();
  const uint2* p = in + (size_t)seg * n;
  const size_t base = (size_t)tile * SORT_TILE;
  const uint32_t count = (uint32_t)((n - base) < (size_t)SORT_TILE ? (n - base) : (size_t)SORT_TILE);
  if ((((uintptr_t)(p + base)) & 15u) == 0) {
    for (uint32_t i = threadIdx.x * 2; i < SORT_TILE; i += SORT_THREADS * 2) {
      if (i + 1 < count) {
        const uint4 v = *reinterpret_cast<const uint4*>(p + base + i);
        atomicAdd(&h[(v.x >> shift) & 0xffu], 1u);
        atomicAdd(&h[(v.z >> shift) & 0xffu], 1u);
      } else if (i < count) {
        atomicAdd(&h[(p[base + i].x >> shift) & 0xffu], 1u);
      }
    }
  } else {
    for (uint32_t i = threadIdx.x; i < count; i += SORT_THREADS) atomicAdd(&h[(p[base + i].x >> shift) & 0xffu], 1u);
  }
  __syncthreads();
  if (threadIdx.x < SORT_RADIX) cnt[sort_cnt_index(seg, threadIdx.x, tile, tiles)] = h[threadIdx.x];
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

// exclusive scan of one value per thread over the first 256 threads (8 warps), two independent sequences at once
__device__ __forceinline__ void scan256x2(uint32_t& a, uint32_t& b, uint32_t (*wsum)[2], int tid) {
  const int lane = tid & 31, warp = tid >> 5;
  uint32_t xa = a, xb = b;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t ya = __shfl_up_sync(0xffffffffu, xa, o), yb = __shfl_up_sync(0xffffffffu, xb, o);
    if (lane >= o) { xa += ya; xb += yb; }
  }
  if (tid < 256 && lane == 31) { wsum[warp][0] = xa; wsum[warp][1] = xb; }
  __syncthreads();
  if (tid < 256) {
    uint32_t oa = 0, ob = 0;
#pragma unroll
    for (int w = 0; w < 8; w++) { oa += (w < warp) ? wsum[w][0] : 0u; ob += (w < warp) ? wsum[w][1] : 0u; }
    a = oa + xa - a;    // exclusive
    b = ob + xb - b;
  }
}

// One tile of 4096 pairs:
//   1. every thread takes 8 pairs (striped: consecutive lanes, consecutive pairs) straight into registers;
//   2. every warp ranks its pairs in index order (stable): a shared-memory mask word per digit collects the lanes with
//      equal digits, the group leader bumps the warp's digit counter;
//   3. per-digit scan over the warps + scan over the digits give each pair its slot in the tile's sorted order; the
//      pairs are written to that slot in shared memory;
//   4. the staged tile is copied out in slot order: consecutive threads write consecutive pairs inside every digit's
//      run (16 pairs = 128 bytes on average).
template <bool LAST>
__global__ void __launch_bounds__(SORT_THREADS, 3)
k_sort_scatter(const uint2* __restrict__ in, uint2* __restrict__ out, uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out,
               size_t n, int shift, const uint32_t* __restrict__ cnt, const uint32_t* __restrict__ rowsum, int tiles) {
  __shared__ __align__(16) uint2 s_pair[SORT_TILE];
  __shared__ uint16_t wcnt[SORT_THREADS / 32][SORT_RADIX];   // per-warp digit counters (<= 4096 each)
  // per-warp "lanes holding this digit" masks (zero between rounds): only alive while the pairs are ranked, before the
  // staging buffer is written, so they share its memory
  uint32_t (*wmask)[SORT_RADIX] = reinterpret_cast<uint32_t (*)[SORT_RADIX]>(s_pair);
  static_assert(sizeof(uint32_t) * (SORT_THREADS / 32) * SORT_RADIX <= sizeof(uint2) * SORT_TILE, "mask words must fit the staging buffer");
  __shared__ uint32_t dstart[SORT_RADIX];   // first slot of digit d in the staged tile
  __shared__ uint32_t gdelta[SORT_RADIX];   // global position of slot i of digit d = gdelta[d] + i
  __shared__ uint32_t wsum[8][2];
  const int seg = blockIdx.y, tile = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint2* pin = in + (size_t)seg * n;
  const size_t tbase = (size_t)tile * SORT_TILE;
  const uint32_t count = (uint32_t)((n - tbase) < (size_t)SORT_TILE ? (n - tbase) : (size_t)SORT_TILE);

  {
    uint32_t* z = reinterpret_cast<uint32_t*>(&wcnt[0][0]);
#pragma unroll
    for (int i = 0; i < (SORT_THREADS / 32) * SORT_RADIX / 2 / SORT_THREADS; i++) z[i * SORT_THREADS + tid] = 0;
    uint32_t* y = reinterpret_cast<uint32_t*>(s_pair);
#pragma unroll
    for (int i = 0; i < (SORT_THREADS / 32) * SORT_RADIX / SORT_THREADS; i++) y[i * SORT_THREADS + tid] = 0;
  }
  // ---- 1. load ----
  uint2 p[SORT_ITEMS];
  const uint32_t wbase = warp * (32 * SORT_ITEMS);
#pragma unroll
  for (int i = 0; i < SORT_ITEMS; i++) {
    const uint32_t idx = wbase + i * 32 + lane;
    // beyond the end of the segment (last tile only): digit 255 by construction, ranked behind every real pair
    p[i] = idx < count ? pin[tbase + idx] : make_uint2(0xffffffffu, 0u);
  }
  __syncthreads();
  // ---- 2. rank ----
  uint32_t off[SORT_ITEMS];
#pragma unroll
  for (int i = 0; i < SORT_ITEMS; i++) {
    const uint32_t d = (p[i].x >> shift) & 0xffu;
    // lanes with the same digit: every lane ORs its bit into the warp's mask word of that digit (shared-memory atomic:
    // ~3 instructions instead of the ~45 of eight ballots; MATCH.ANY costs one round per DISTINCT value, ~30 here)
    atomicOr(&wmask[warp][d], 1u << lane);
    __syncwarp();
    const uint32_t m = wmask[warp][d];
    __syncwarp();
    const int leader = __ffs(m) - 1;
    const uint32_t before = __popc(m & ((1u << lane) - 1u));
    uint32_t old = 0;
    if (lane == leader) {
      old = wcnt[warp][d];
      wcnt[warp][d] = (uint16_t)(old + __popc(m));
      wmask[warp][d] = 0;          // ready for the next round (the __syncwarp below orders it)
    }
    old = __shfl_sync(0xffffffffu, old, leader);
    off[i] = old + before;
    __syncwarp();
  }
  __syncthreads();
  // ---- 3. slots ----
  uint32_t tot = 0, gb = 0;
  if (tid < SORT_RADIX) {
#pragma unroll
    for (int w = 0; w < SORT_THREADS / 32; w++) {
      const uint32_t t = wcnt[w][tid];
      wcnt[w][tid] = (uint16_t)tot;
      tot += t;
    }
    gb = rowsum[(size_t)seg * SORT_RADIX + tid];
  }
  {
    uint32_t slot0 = tot, base = gb;
    scan256x2(slot0, base, wsum, tid);         // (contains a __syncthreads)
    if (tid < SORT_RADIX) {
      dstart[tid] = slot0;
      gdelta[tid] = base + cnt[sort_cnt_index(seg, tid, tile, tiles)] - slot0;
    }
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < SORT_ITEMS; i++) {
    const uint32_t d = (p[i].x >> shift) & 0xffu;
    const uint32_t slot = dstart[d] + wcnt[warp][d] + off[i];
    s_pair[slot] = p[i];          // padding pairs land at slots >= count (digit 255, last ranks)
  }
  __syncthreads();
  // ---- 4. copy out in slot order ----
  const size_t obase = (size_t)seg * n;
#pragma unroll 2
  for (uint32_t i = tid; i < count; i += SORT_THREADS) {
    const uint2 q = s_pair[i];
    const uint32_t d = (q.x >> shift) & 0xffu;
    const uint32_t pos = gdelta[d] + i;
    if (LAST) {
      keys_out[obase + pos] = q.x;
      vals_out[obase + pos] = q.y;
    } else {
      out[obase + pos] = q;
    }
  }
}

int sort_pairs(cudaStream_t s, uint2* a, uint2* b, uint32_t* keys_out, uint32_t* vals_out, size_t n, int nseg, int key_bits,
               uint32_t* cnt, uint32_t* rowsum, int tiles) {
  const dim3 g(tiles, nseg);
  int launches = 0;
  int passes = (key_bits + 7) / 8;
  if (passes < 1) passes = 1;
  for (int pass = 0; pass < passes; pass++) {
    const bool last = pass == passes - 1;
    k_sort_hist<<<g, SORT_THREADS, 0, s>>>(a, n, 8 * pass, cnt, tiles);
    k_sort_rowscan<<<dim3(SORT_RADIX, nseg), 256, 0, s>>>(cnt, rowsum, tiles);
    if (last) k_sort_scatter<true><<<g, SORT_THREADS, 0, s>>>(a, b, keys_out, vals_out, n, 8 * pass, cnt, rowsum, tiles);
    else k_sort_scatter<false><<<g, SORT_THREADS, 0, s>>>(a, b, nullptr, nullptr, n, 8 * pass, cnt, rowsum, tiles);
    launches += 3;
    uint2* t = a; a = b; b = t;
  }
  return launches;
}

}  // namespace zk
