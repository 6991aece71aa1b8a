// Bn254G2: affine pre-reduction tree + record accumulation
#include "kernels_aff.cuh"
namespace zk {
ZK_INSTANTIATE_AFF(Bn254G2)
}
