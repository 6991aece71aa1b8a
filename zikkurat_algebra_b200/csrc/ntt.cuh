// Fr NTT launcher (ntt.cu), see there for the algorithm.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace zk {

// N = 2^m elements of 8 x u32 (canonical Montgomery Fr); all pointers are device memory of the current device:
// d_gen 8 words, d_src / d_tmp / d_dst N*8 words (distinct buffers), d_table (N/2 + 1)*8 words.
template <class F>
void ntt_device(cudaStream_t s, int m, const uint32_t* d_gen, const uint32_t* d_src, uint32_t* d_tmp, uint32_t* d_dst,
                uint32_t* d_table, int inverse);

// table[i] = gen^i (Montgomery) for i < half = N/2 (N = 2^m), table[half] = N^-1; (half + 1) * 8 words
template <class F>
void ntt_build_table(cudaStream_t s, const uint32_t* d_gen, size_t half, int m, uint32_t* d_table);

}  // namespace zk
