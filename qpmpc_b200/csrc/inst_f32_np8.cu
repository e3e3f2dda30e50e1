// Explicit instantiations of the warp kernels: float, NP = 8.
#define QPMPC_INSTANTIATE
#include "mpc_launch.cuh"

namespace qpmpc {
QPMPC_INSTANTIATE_VARIANT(float, 8, 2, true)
QPMPC_INSTANTIATE_VARIANT(float, 8, 4, true)
QPMPC_INSTANTIATE_PAIRED(float, 8)
}  // namespace qpmpc
