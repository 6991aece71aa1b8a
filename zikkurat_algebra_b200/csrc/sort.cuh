// Segmented LSD radix sort of (bucket key, point index) pairs, 8 bits per pass.
//
// This is the explicit form of the reference's implicit "scatter every point into SUMS[e]"
//   lib/cbits/curves/g1/proj/bn128_G1_proj.c:549-561
// One segment = one Pippenger window (of one MSM of a batch); segments are sorted independently,
// all in the same launches (blockIdx.y = segment).
//
// Pairs travel PACKED (uint2 {key, value}: one 8-byte access per pair, a whole 128-byte line per 16-pair digit
// run); the last pass writes the two separate arrays the accumulation kernels read.  Per pass, three kernels
// (deterministic, stable):
//   k_sort_hist     tile digit histogram                                  -> cnt[seg][digit][tile]
//   k_sort_rowscan  exclusive scan of the tile counters along the tiles   -> cnt in place, rowsum[seg][digit]
//   k_sort_scatter  pairs straight into registers, stable in-tile rank (per-warp digit masks in shared memory), staged
//                   in shared memory in sorted order, copied out coalesced per digit run
// HBM traffic per pass and pair: 8 B (hist) + 8 B read + 8 B written.
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>

namespace zk {

constexpr int SORT_THREADS = 512;
constexpr int SORT_ITEMS = 8;                           // pairs per thread
constexpr int SORT_TILE = SORT_THREADS * SORT_ITEMS;    // 4096 pairs per CTA
constexpr int SORT_TILE_LOG = 12;
constexpr int SORT_RADIX = 256;

// counter of (segment, digit, tile)
__host__ __device__ inline size_t sort_cnt_index(size_t seg, uint32_t digit, size_t tile, int tiles) {
  return (seg * SORT_RADIX + digit) * (size_t)tiles + tile;
}
inline size_t sort_cnt_bytes(int nseg, int tiles) { return (size_t)nseg * SORT_RADIX * (size_t)tiles * 4; }

// All passes for keys of `key_bits` bits: pairs in `a`, `b` is the other packed buffer; result in keys_out / vals_out
// (segment-major, n per segment).  cnt: sort_cnt_bytes.  Returns the number of kernels launched.
int sort_pairs(cudaStream_t s, uint2* a, uint2* b, uint32_t* keys_out, uint32_t* vals_out, size_t n, int nseg, int key_bits,
               uint32_t* cnt, uint32_t* rowsum, int tiles);

}  // namespace zk
