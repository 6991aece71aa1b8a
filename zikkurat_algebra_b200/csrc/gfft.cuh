// FFT of G1 group elements (gfft.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace zk {

// N = 2^m projective (jac = 0) or Jacobian (jac = 1) points (3L words each) at d_src -> normalised projective points at d_dst; d_work holds N XYZZ
// records (4L words each), d_table (N/2 + 1) * 8 words, d_gen 8 words (Montgomery Fr generator of the order-N subgroup)
template <class C>
void gfft_device(cudaStream_t s, int m, const uint32_t* d_gen, const uint32_t* d_src, void* d_work, uint32_t* d_table,
                 uint32_t* d_dst, int inverse, int jac, int glv);
// glv != 0: twiddle products through the curve endomorphism where the curve has one (G1; subgroup points, see glv.cuh)

}  // namespace zk
