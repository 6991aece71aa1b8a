// Element-wise self-test of the device field / group primitives (selftest.cu); op codes shared with the host layer
// and mirrored in zikkurat_algebra_b200/__init__.py.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace zk {

enum SelftestFieldOp : int {
  ZKT_MUL = 0, ZKT_SQR = 1, ZKT_MUL2 = 2, ZKT_ADD = 3, ZKT_SUB = 4, ZKT_NEG = 5, ZKT_INV = 6,
  ZKT_MUL_CALL = 7, ZKT_SQR_CALL = 8, ZKT_MUL2_CALL = 9, ZKT_DBL = 10, ZKT_FROM_MONT = 11,
  ZKT_PAIR_FIRST = 12, ZKT_PAIR_SECOND = 13   // a*b resp. a*c out of the paired product fe_mul_pair_call(a, b, c)
};
enum SelftestGroupOp : int {
  ZKT_G_MADD = 0, ZKT_G_MADD_CALLS = 1, ZKT_G_ADD = 2, ZKT_G_ADD_CALLS = 3, ZKT_G_DBL = 4, ZKT_G_DBL_AFFINE = 5
};

template <class P>
void launch_selftest_field(cudaStream_t s, int op, size_t n, const uint32_t* a, const uint32_t* b, const uint32_t* c,
                           const uint32_t* d, uint32_t* out);
template <class C>
void launch_selftest_group(cudaStream_t s, int op, size_t n, const uint32_t* p1, const uint32_t* z1, const uint32_t* p2,
                           const uint32_t* z2, uint32_t* out);

}  // namespace zk
