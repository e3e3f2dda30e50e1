// Explicit instantiations of the warp kernels: double, NP = 16.
#define QPMPC_INSTANTIATE
#include "mpc_launch.cuh"

namespace qpmpc {
QPMPC_INSTANTIATE_VARIANT(double, 16, 2, true)
QPMPC_INSTANTIATE_PDIP(16, 2)
QPMPC_INSTANTIATE_VARIANT(double, 16, 4, false)
QPMPC_INSTANTIATE_PDIP(16, 4)
QPMPC_INSTANTIATE_PAIRED(double, 16)
}  // namespace qpmpc
