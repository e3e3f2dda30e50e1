// pdip_emu.cpp -- host build of pdip_core() (qpmpc_b200/csrc/mpc_pdip.cuh) on
// the fiber warp emulator.  TEST INFRASTRUCTURE ONLY: lets the CPU test-suite
// run the device source of the interior-point phases, lane by lane, against
// the NumPy model and the exact oracle.  Built by tests/test_pdip_emu.py:
//   g++ -O1 -std=c++17 -shared -fPIC -DQPMPC_HOST_EMU -I<repo> pdip_emu.cpp
#include "warp_emu.h"

#include "qpmpc_b200/csrc/mpc_pdip.cuh"

#include <vector>

namespace {

// `count` QPs (count <= 32 / NP) side by side in one emulated warp, exactly as
// the kernel lays them out: P, G by columns with the kernel's leading
// dimensions, padding variables with P = identity.
template <typename T, int NP, int MR>
int run(int count, int n, int m, const double *P, const double *q, const double *G, const double *h, int max_iter,
        double tol, int polish, double *U, double *Z, int *status, int *iters) {
    using L = qpmpc::PdipLay<T, NP, MR>;
    constexpr int IPW = 32 / NP;
    if (count < 1 || count > IPW || n > NP || m > L::MP) return -1;
    const int stride = L::fixed + L::szG;
    const T junk = std::numeric_limits<T>::quiet_NaN();  // what an unwritten region may hold
    std::vector<T> smem((size_t)IPW * stride, junk);
    for (int g = 0; g < IPW; ++g) {
        const int src = g < count ? g : 0;  // tail slots read instance 0, like the kernel
        T *wk = smem.data() + (size_t)g * stride;
        T *hs = wk + L::oH, *Pc = wk + L::oP, *Gc = wk + L::fixed;
        for (int r = 0; r < m; ++r) hs[r] = (T)h[(size_t)src * m + r];
        for (int c = 0; c < NP; ++c)
            for (int r = 0; r < NP; ++r)
                Pc[c * L::LDL + r] = (r < n && c < n) ? (T)P[((size_t)src * n + r) * n + c] : (r == c ? T(1) : T(0));
        for (int c = 0; c < NP; ++c)
            for (int r = 0; r < m; ++r) Gc[c * L::LDG + r] = c < n ? (T)G[((size_t)src * m + r) * n + c] : T(0);
    }
    emu::run_warp([&](int lane) {
        const int sub = lane / NP, l = lane % NP;
        T *wk = smem.data() + (size_t)sub * stride;
        T *xs = wk + L::oV, *dv = xs + NP, *wv = wk + L::oW, *tv = wv + L::MP;
        const bool valid = sub < count;
        const int src = valid ? sub : 0;
        const T qj = l < n ? (T)q[(size_t)src * n + l] : T(0);
        T x, z[MR];
        int st, it;
        qpmpc::pdip_core<T, NP, MR>(wk + L::oP, qj, wk + L::fixed, wk + L::oH, wk + L::oL, xs, dv, wv, tv, m, l, valid,
                                    max_iter, (T)tol, polish != 0, x, z, st, it);
        if (!valid) return;
        if (l < n) U[(size_t)sub * n + l] = (double)x;
        for (int s = 0; s < MR; ++s)
            if (l + s * NP < m) Z[(size_t)sub * m + l + s * NP] = (double)z[s];
        if (l == 0) {
            status[sub] = st;
            iters[sub] = it;
        }
    });
    return 0;
}

}  // namespace

extern "C" int pdip_emu_solve(int dtype, int np, int mr, int count, int n, int m, const double *P, const double *q,
                              const double *G, const double *h, int max_iter, double tol, int polish, double *U,
                              double *Z, int *status, int *iters) {
#define CASE(T, NP, MR)                   \
    if (np == NP && mr == MR)             \
        return run<T, NP, MR>(count, n, m, P, q, G, h, max_iter, tol, polish, U, Z, status, iters);
    if (dtype == 0) {
        CASE(double, 8, 2)
        CASE(double, 8, 4)
        CASE(double, 16, 2)
        CASE(double, 16, 4)
        CASE(double, 32, 2)
        CASE(double, 32, 4)
    } else {
        CASE(float, 8, 2)
        CASE(float, 16, 2)
        CASE(float, 32, 2)
    }
#undef CASE
    return -2;
}
