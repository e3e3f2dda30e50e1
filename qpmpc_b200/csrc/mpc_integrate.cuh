// mpc_integrate.cuh -- batched MPCProblem.integrate (what Plan.states calls).
#pragma once

#include "mpc_common.cuh"

namespace qpmpc {

// ---------------------------------------------------------------------------
// X_{k+1} = A_k X_k + B_k U_k, one thread per instance:
// MPCProblem.integrate (qpmpc/mpc_problem.py:316-335).
// ---------------------------------------------------------------------------
struct IntegrateParams {
    int batch, N, nx, nu;
    const void *A, *B, *x0, *U;
    long long bA, bB, bx0;  // batch strides (elements), 0 if shared
    int sA, sB;             // step strides (elements), 0 if LTI
    void *X;
};

template <typename T>
__global__ void mpc_integrate_kernel(const IntegrateParams p) {
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= p.batch) return;
    const int nx = p.nx, nu = p.nu, N = p.N;
    const T *A = static_cast<const T *>(p.A) + b * p.bA;
    const T *B = static_cast<const T *>(p.B) + b * p.bB;
    const T *x0 = static_cast<const T *>(p.x0) + b * p.bx0;
    const T *U = static_cast<const T *>(p.U) + b * (long long)N * nu;
    T *X = static_cast<T *>(p.X) + b * (long long)(N + 1) * nx;
    for (int t = 0; t < nx; ++t) X[t] = x0[t];
    for (int k = 0; k < N; ++k) {
        const T *Ak = A + (long long)k * p.sA, *Bk = B + (long long)k * p.sB;
        const T *xk = X + (long long)k * nx;
        T *xn = X + (long long)(k + 1) * nx;
        for (int t = 0; t < nx; ++t) {
            T acc = T(0);
            for (int s = 0; s < nx; ++s) acc += Ak[t * nx + s] * xk[s];
            for (int s = 0; s < nu; ++s) acc += Bk[t * nu + s] * U[k * nu + s];
            xn[t] = acc;
        }
    }
}

}  // namespace qpmpc
