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
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>

namespace zk {

constexpr int SORT_THREADS = 512;
constexpr int SORT_ITEMS = 8;                           // keys per thread
constexpr int SORT_TILE = SORT_THREADS * SORT_ITEMS;    // 4096 pairs per CTA
constexpr int SORT_RADIX = 256;

// one 8-bit pass over all segments: (keys_in, vals_in) -> (keys_out, vals_out), stable
void sort_pass(cudaStream_t s, const uint32_t* keys_in, const uint32_t* vals_in, uint32_t* keys_out, uint32_t* vals_out,
               size_t n, int nseg, int shift, uint32_t* cnt, uint32_t* rowsum, int tiles);

}  // namespace zk
