// Explicit instantiations of the warp kernels: float, NP = 16.
#define QPMPC_INSTANTIATE
#include "mpc_launch.cuh"

namespace qpmpc {
QPMPC_INSTANTIATE_VARIANT(float, 16, 2, true)
QPMPC_INSTANTIATE_VARIANT(float, 16, 4, false)
QPMPC_INSTANTIATE_PAIRED(float, 16)
}  // namespace qpmpc
