// Bn254G2: recode + bucket accumulation kernels (G2, coordinates in Fp2)
#include "kernels_acc.cuh"
namespace zk {
ZK_INSTANTIATE_ACC(Bn254G2)
}
