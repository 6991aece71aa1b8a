// Bls12381: bucket reduction + tail kernels
#include "kernels_red.cuh"
namespace zk {
ZK_INSTANTIATE_RED(Bls12381)
}
