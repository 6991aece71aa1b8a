// Bls12381: recode + bucket accumulation kernels
#include "kernels_acc.cuh"
namespace zk {
ZK_INSTANTIATE_ACC(Bls12381)
}
