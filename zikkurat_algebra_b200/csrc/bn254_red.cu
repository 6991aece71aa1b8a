// Bn254: bucket reduction + tail kernels
#include "kernels_red.cuh"
namespace zk {
ZK_INSTANTIATE_RED(Bn254)
}
