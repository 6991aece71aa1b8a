// Bn254: recode + bucket accumulation kernels
#include "kernels_acc.cuh"
namespace zk {
ZK_INSTANTIATE_ACC(Bn254)
}
