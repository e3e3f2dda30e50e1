// Explicit instantiations of the warp kernels: float, NP = 32.
#define QPMPC_INSTANTIATE
#include "mpc_launch.cuh"

namespace qpmpc {
QPMPC_INSTANTIATE_VARIANT(float, 32, 2, false)
QPMPC_INSTANTIATE_VARIANT(float, 32, 4, false)
QPMPC_INSTANTIATE_PAIRED(float, 32)
}  // namespace qpmpc
