// Explicit instantiations of the warp kernels: double, NP = 32.
#define QPMPC_INSTANTIATE
#include "mpc_launch.cuh"

namespace qpmpc {
QPMPC_INSTANTIATE_VARIANT(double, 32, 2, false)
QPMPC_INSTANTIATE_PDIP(32, 2)
QPMPC_INSTANTIATE_VARIANT(double, 32, 4, false)
QPMPC_INSTANTIATE_PDIP(32, 4)
QPMPC_INSTANTIATE_PAIRED(double, 32)
}  // namespace qpmpc
