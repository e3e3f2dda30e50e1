// Explicit instantiations of the low-rank-Hessian kernel (mpc_lr_kernel.cuh): double, NP = 32 / 64.
#include "mpc_launch.cuh"
#include "mpc_lr_kernel.cuh"

namespace qpmpc {

template <typename T, int NP>
int launch_solve_lr(SolveParams p, cudaStream_t stream) {
    const size_t smem = lr_layout_smem<T, NP>(&p);
    if (smem > 227 * 1024) return QPMPC_B200_ESHAPE;
    if (p.batch == 0) return 0;
    auto launch = [&](auto kern) {
        cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (err != cudaSuccess) return (int)err;
        kern<<<p.batch, NP, smem, stream>>>(p);
        count_launch();
        return (int)cudaGetLastError();
    };
    switch (p.nx) {
        case 2: return launch(mpc_solve_lr_kernel<T, NP, 2>);
        case 3: return launch(mpc_solve_lr_kernel<T, NP, 3>);
        case 4: return launch(mpc_solve_lr_kernel<T, NP, 4>);
    }
    return QPMPC_B200_ESHAPE;
}

template int launch_solve_lr<double, 16>(SolveParams, cudaStream_t);
template int launch_solve_lr<double, 32>(SolveParams, cudaStream_t);
template int launch_solve_lr<double, 64>(SolveParams, cudaStream_t);

}  // namespace qpmpc
