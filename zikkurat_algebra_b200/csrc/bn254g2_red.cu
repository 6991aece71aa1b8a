// Bn254G2: bucket reduction + tail kernels (G2, coordinates in Fp2)
#include "kernels_red.cuh"
namespace zk {
ZK_INSTANTIATE_RED(Bn254G2)
}
