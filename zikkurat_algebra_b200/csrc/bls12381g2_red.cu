// Bls12381G2: bucket reduction + tail kernels (G2, coordinates in Fp2)
#include "kernels_red.cuh"
namespace zk {
ZK_INSTANTIATE_RED(Bls12381G2)
}
