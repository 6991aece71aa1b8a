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
//   k_sort_scatter  stable in-tile rank (warp match) + global offset, writes the permuted pairs
// HBM traffic per pass and pair: 4 B (hist) + 8 B read + 8 B written.
#include "sort.cuh"

namespace zk {

__global__ void __launch_bounds__(SORT_THREADS)
k_sort_hist(const uint32_t* __restrict__ keys, size_t n, int shift, uint32_t* __restrict__ cnt, int tiles) {
  __shared__ uint32_t h[SORT_RADIX];
  const int seg = blockIdx.y, tile = blockIdx.x;
  h[threadIdx.x] = 0;
  __syncthreads();
  const uint32_t* k = keys + (size_t)seg * n;
  size_t base = (size_t)tile * SORT_TILE;
#pragma unroll
  for (int i = 0; i < SORT_ITEMS; i++) {
    size_t idx = base + (size_t)i * SORT_THREADS + threadIdx.x;
    if (idx < n) atomicAdd(&h[(k[idx] >> shift) & 0xffu], 1u);
  }
  __syncthreads();
  cnt[((size_t)seg * SORT_RADIX + threadIdx.x) * tiles + tile] = h[threadIdx.x];
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

__global__ void __launch_bounds__(SORT_THREADS)
k_sort_scatter(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
               uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out, size_t n, int shift,
               const uint32_t* __restrict__ cnt, const uint32_t* __restrict__ rowsum, int tiles) {
  __shared__ uint32_t wcnt[SORT_THREADS / 32][SORT_RADIX + 1];
  __shared__ uint32_t dbase[SORT_RADIX];
  __shared__ uint32_t scan_tmp[SORT_RADIX];
  const int seg = blockIdx.y, tile = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (SORT_THREADS / 32) * (SORT_RADIX + 1); i += SORT_THREADS) (&wcnt[0][0])[i] = 0;
  // exclusive scan of the 256 digit totals of this segment (Hillis-Steele in shared memory)
  uint32_t mine = rowsum[(size_t)seg * SORT_RADIX + threadIdx.x];
  scan_tmp[threadIdx.x] = mine;
  __syncthreads();
  for (int o = 1; o < SORT_RADIX; o <<= 1) {
    uint32_t y = (threadIdx.x >= o) ? scan_tmp[threadIdx.x - o] : 0u;
    __syncthreads();
    scan_tmp[threadIdx.x] += y;
    __syncthreads();
  }
  dbase[threadIdx.x] = scan_tmp[threadIdx.x] - mine + cnt[((size_t)seg * SORT_RADIX + threadIdx.x) * tiles + tile];

  const uint32_t* kin = keys_in + (size_t)seg * n;
  const uint32_t* vin = vals_in + (size_t)seg * n;
  size_t wbase = (size_t)tile * SORT_TILE + (size_t)warp * (32 * SORT_ITEMS);
  uint32_t key[SORT_ITEMS], val[SORT_ITEMS], off[SORT_ITEMS];
#pragma unroll
  for (int i = 0; i < SORT_ITEMS; i++) {
    size_t idx = wbase + (size_t)i * 32 + lane;
    bool ok = idx < n;
    key[i] = ok ? kin[idx] : 0u;
    val[i] = ok ? vin[idx] : 0u;
  }
#pragma unroll
  for (int i = 0; i < SORT_ITEMS; i++) {
    size_t idx = wbase + (size_t)i * 32 + lane;
    uint32_t d = (idx < n) ? ((key[i] >> shift) & 0xffu) : (uint32_t)SORT_RADIX;  // tail lanes share a dummy bin
    uint32_t peers = __match_any_sync(0xffffffffu, d);
    int leader = __ffs(peers) - 1;
    uint32_t before = __popc(peers & ((1u << lane) - 1u));
    uint32_t old = 0;
    if (lane == leader) {
      old = wcnt[warp][d];
      wcnt[warp][d] = old + __popc(peers);
    }
    old = __shfl_sync(0xffffffffu, old, leader);
    off[i] = old + before;
    __syncwarp();
  }
  __syncthreads();
  {  // per digit: exclusive scan across the warps of this CTA
    uint32_t run = 0;
#pragma unroll
    for (int w = 0; w < SORT_THREADS / 32; w++) {
      uint32_t t = wcnt[w][threadIdx.x];
      wcnt[w][threadIdx.x] = run;
      run += t;
    }
  }
  __syncthreads();
  uint32_t* kout = keys_out + (size_t)seg * n;
  uint32_t* vout = vals_out + (size_t)seg * n;
#pragma unroll
  for (int i = 0; i < SORT_ITEMS; i++) {
    size_t idx = wbase + (size_t)i * 32 + lane;
    if (idx < n) {
      uint32_t d = (key[i] >> shift) & 0xffu;
      size_t pos = (size_t)dbase[d] + wcnt[warp][d] + off[i];
      kout[pos] = key[i];
      vout[pos] = val[i];
    }
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
