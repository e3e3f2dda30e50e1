// Explicit instantiations of the warp kernels: double, NP = 8.
#define QPMPC_INSTANTIATE
#include "mpc_launch.cuh"

namespace qpmpc {
QPMPC_INSTANTIATE_VARIANT(double, 8, 2, true)
QPMPC_INSTANTIATE_PDIP(8, 2)
QPMPC_INSTANTIATE_VARIANT(double, 8, 4, true)
QPMPC_INSTANTIATE_PDIP(8, 4)
QPMPC_INSTANTIATE_PAIRED(double, 8)
}  // namespace qpmpc
