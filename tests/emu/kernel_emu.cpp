// kernel_emu.cpp -- host build of the engine's warp kernels on the fiber
// emulator (warp_emu.h).  TEST INFRASTRUCTURE ONLY.
//
// emu_solve() takes the C ABI's own descriptor / operand / output structs
// (include/qpmpc_b200.h, HOST pointers) and does what qpmpc_b200_solve does --
// check_desc, fill_params, pick_variant, layout_smem from the product's headers
// -- but "launches" mpc_solve_kernel / mpc_pdip_kernel (the device SOURCE,
// compiled for the host) CTA by CTA on fibers.  emu_condense() does the same
// for mpc_condense_kernel.  tests/test_kernel_emu.py compares the results with
// the oracle; nothing here is measured or shipped.
//
//   g++ -O1 -std=c++17 -shared -fPIC -DQPMPC_HOST_EMU -Wno-unknown-pragmas \
//       -I<repo> tests/emu/kernel_emu.cpp -o tests/emu/libkernel_emu.so
#include "warp_emu.h"

#include <cstdlib>

#include "qpmpc_b200/csrc/mpc_host_params.h"
#include "qpmpc_b200/csrc/mpc_pdip.cuh"  // brings mpc_common.cuh and mpc_kernels.cuh
#include "qpmpc_b200/csrc/mpc_cta_kernel.cuh"
#include "qpmpc_b200/csrc/mpc_lr_kernel.cuh"
#include "qpmpc_b200/csrc/mpc_factor.cuh"
#include "qpmpc_b200/csrc/mpc_integrate.cuh"
#include "qpmpc_b200/csrc/mpc_plant.cuh"

namespace qpmpc {
alignas(16) unsigned char smem_raw[256 * 1024];  // what `extern __shared__` names in the kernels
int env_int(const char *name, int dflt) {
    const char *v = std::getenv(name);
    return (v && *v) ? std::atoi(v) : dflt;
}
}  // namespace qpmpc

using namespace qpmpc;

namespace {

int g_smem_overrun = 0;  // CTAs that wrote past their dynamic shared-memory request

// one grid: CTAs one after the other, each on a NaN-poisoned shared-memory image
// followed by a canary (a store past the requested size is an error on the device)
template <typename K>
void launch(int grid, int threads, size_t smem, K kernel) {
    const uint64_t poison = 0x7ff8dead7ff8deadull;  // NaN as double and as two floats
    const size_t lo = (smem + 7) / 8 * 8;
    for (int b = 0; b < grid; ++b) {
        for (size_t i = 0; i + 8 <= sizeof(smem_raw); i += 8) memcpy(smem_raw + i, &poison, 8);
        emu::run_cta(threads, emu::Idx{(unsigned)b, 0, 0}, emu::Idx{(unsigned)grid, 1, 1}, kernel);
        for (size_t i = lo; i + 8 <= sizeof(smem_raw); i += 8)
            if (memcmp(smem_raw + i, &poison, 8) != 0) {
                ++g_smem_overrun;
                break;
            }
    }
}

template <typename T, int NP, int MR, bool MREG, bool RS, bool PAIRED = false>
int solve_variant(SolveParams p, int wpc) {
    using L = Lay<T, NP, MR, MREG, RS, PAIRED>;
    constexpr int IPW = 32 / NP;
    size_t smem = 0;
    for (;; --wpc) {
        smem = layout_smem<T>(&p, L::fixed, L::szG, NP, IPW * wpc, MREG);
        if (smem <= 227 * 1024 || wpc == 1) break;
    }
    if (smem > 227 * 1024) return QPMPC_B200_ESHAPE;
    const int ipc = IPW * wpc;
    launch((p.batch + ipc - 1) / ipc, wpc * 32, smem, [&]() { mpc_solve_kernel<T, NP, MR, MREG, RS, PAIRED>(p); });
    return 0;
}

// launch_solve (mpc_launch.cuh): the NP = 16 register-resident kernel has two builds
template <typename T, int NP, int MR, bool MREG>
int solve(const SolveParams &p, int wpc) {
    if constexpr (NP == 16 && MREG) {
        const int rs = env_int("QPMPC_B200_ROWS_SMEM", -1);
        if (rs > 0 || (rs < 0 && !p.has_wx)) return solve_variant<T, NP, MR, MREG, true>(p, wpc);
    }
    return solve_variant<T, NP, MR, MREG, false>(p, wpc);
}

template <typename T, int NP, int MR>
int pdip(SolveParams p, int polish, int wpc) {
    using L = PdipLay<T, NP, MR>;
    constexpr int IPW = 32 / NP;
    size_t smem = 0;
    for (;; --wpc) {
        smem = layout_smem<T>(&p, L::fixed, L::szG, NP, IPW * wpc, false, true);
        if (smem <= 227 * 1024 || wpc == 1) break;
    }
    if (smem > 227 * 1024) return QPMPC_B200_ESHAPE;
    const int ipc = IPW * wpc;
    if (env_int("QPMPC_B200_PDIP_SOLVE", 0) != 0)
        launch((p.batch + ipc - 1) / ipc, wpc * 32, smem, [&]() { mpc_pdip_kernel<T, NP, MR, true>(p, polish); });
    else
        launch((p.batch + ipc - 1) / ipc, wpc * 32, smem, [&]() { mpc_pdip_kernel<T, NP, MR, false>(p, polish); });
    return 0;
}

template <typename T, int NP, int MR>
int condense(SolveParams p, int wpc) {
    using L = Lay<T, NP, MR, false>;
    constexpr int IPW = 32 / NP;
    size_t smem = 0;
    for (;; --wpc) {
        smem = layout_smem<T>(&p, L::fixed, L::szG, NP, IPW * wpc, false);
        if (smem <= 227 * 1024 || wpc == 1) break;
    }
    if (smem > 227 * 1024) return QPMPC_B200_ESHAPE;
    const int ipc = IPW * wpc, grid = (p.batch + ipc - 1) / ipc;
    SolveParams a = p, b = p;
    a.Phi = a.Psi = a.phi_last = a.psi_last = nullptr;
    b.P = b.q = b.G = b.h = nullptr;
    if (p.P || p.q || p.G || p.h) launch(grid, wpc * 32, smem, [&]() { mpc_condense_kernel<T, NP, MR, false>(a); });
    if (p.Phi || p.Psi || p.phi_last || p.psi_last)
        launch(grid, wpc * 32, smem, [&]() { mpc_condense_kernel<T, NP, MR, true>(b); });
    return 0;
}


// launch_solve_cta / launch_condense_cta (qpmpc_b200.cu): one CTA per instance, or -- matrices
// beyond 227 KB, or QPMPC_B200_CTA_WORKSPACE=1 -- a bounded grid (3 CTAs here, so that the stride
// over the batch is exercised) with the matrices in a workspace outside shared memory
template <typename T, typename K>
int run_cta(SolveParams p, int threads, K kernel) {
    const CtaLay L = cta_layout(p.n, p.m, p.nx, (int)sizeof(T));
    size_t smem = (size_t)L.total * sizeof(T);
    int grid = p.batch;
    std::vector<T> ws;
    if (smem > 227 * 1024 || env_int("QPMPC_B200_CTA_WORKSPACE", 0) != 0) {
        smem = (size_t)(L.total - L.oV) * sizeof(T);
        if (smem > 227 * 1024 || p.m > 4096 || p.n > 512) return QPMPC_B200_ESHAPE;
        grid = p.batch < 3 ? p.batch : 3;
        ws.assign((size_t)grid * L.oV, std::numeric_limits<T>::quiet_NaN());
        p.workspace = ws.data();
    }
    launch(grid, threads, smem, [&]() { kernel(p); });
    return 0;
}
template <typename T>
int solve_cta(SolveParams p, int threads) {
    p.toeplitz = env_int("QPMPC_B200_NO_TOEPLITZ", 0) == 0;
    return run_cta<T>(p, threads, [](const SolveParams &q) { mpc_solve_cta_kernel<T>(q); });
}
template <typename T>
int condense_cta(const SolveParams &p) {
    return run_cta<T>(p, 256, [](const SolveParams &q) { mpc_condense_cta_kernel<T>(q); });
}

// ---- pdip_core() on explicit QPs (P, q, G, h given), one emulated warp ------
// `count` QPs (count <= 32 / NP) side by side, laid out as the kernel lays
// them out: P, G by columns with the kernel's leading dimensions, padding
// variables with P = identity.  Lets the tests reach cases the MPC front-end
// never produces (infeasible rows, m = 0) and the float instantiation.
template <typename T, int NP, int MR, bool LS>
int pdip_core_run(int count, int n, int m, const double *P, const double *q, const double *G, const double *h,
                  int max_iter, double tol, int polish, double pmin, double *U, double *Z, int *status, int *iters) {
    using L = PdipLay<T, NP, MR>;
    constexpr int IPW = 32 / NP;
    if (count < 1 || count > IPW || n > NP || m > L::MP) return -1;
    const int stride = L::fixed + L::szG;
    std::vector<T> smem((size_t)IPW * stride, std::numeric_limits<T>::quiet_NaN());
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
    emu::run_cta(32, emu::Idx{0, 0, 0}, emu::Idx{1, 1, 1}, [&]() {
        const int lane = emu::lane_id();
        const int sub = lane / NP, l = lane % NP;
        T *wk = smem.data() + (size_t)sub * stride;
        T *xs = wk + L::oV, *dv = xs + NP, *wv = wk + L::oW, *tv = wv + L::MP;
        const bool valid = sub < count;
        const int src = valid ? sub : 0;
        const T qj = l < n ? (T)q[(size_t)src * n + l] : T(0);
        T x, z[MR];
        int st, it;
        pdip_core<T, NP, MR, LS>(wk + L::oP, qj, wk + L::fixed, wk + L::oH, wk + L::oL, xs, dv, wv, tv, m, l, valid,
                                 max_iter, (T)tol, polish != 0, (T)pmin, x, z, st, it);
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

// launch_solve_lr (inst_lr.cu): one CTA of NP threads per instance
template <typename T, int NP>
int solve_lr(SolveParams p) {
    const size_t smem = lr_layout_smem<T, NP>(&p);
    if (smem > 227 * 1024) return QPMPC_B200_ESHAPE;
    switch (p.nx) {
        case 2: launch(p.batch, NP, smem, [&]() { mpc_solve_lr_kernel<T, NP, 2>(p); }); return 0;
        case 3: launch(p.batch, NP, smem, [&]() { mpc_solve_lr_kernel<T, NP, 3>(p); }); return 0;
        case 4: launch(p.batch, NP, smem, [&]() { mpc_solve_lr_kernel<T, NP, 4>(p); }); return 0;
    }
    return QPMPC_B200_ESHAPE;
}

// launch_solve_pre (mpc_launch.cuh): shared-model fast path
template <typename T, int NP>
int solve_pre(SolveParams p, int wpc) {
    using L = Lay<T, NP, 1, true, false, true, true>;
    constexpr int IPW = 32 / NP;
    for (int o : {OP_A, OP_B, OP_C, OP_D}) p.op[o] = OperandView{nullptr, 0, 0, 0, 0};
    p.toeplitz = 0;
    p.inst_stride = (L::fixed + 3) / 4 * 4;
    p.gt_off = p.g_off = p.scr_off = 0;
    const FactorLay F = factor_layout(NP, p.nx, p.N, p.q_wx != 0);
    if (wpc <= 0) wpc = 8;
    const int ipc = IPW * wpc;
    int off = 0;
    p.present_mask = 0;
    for (int o = 0; o < OP_COUNT; ++o) {
        OperandView &v = p.op[o];
        if (!v.ptr) continue;
        p.present_mask |= 1 << o;
        v.smem_off = off;
        off += (v.sz * (v.per_instance ? ipc : 1) + 3) / 4 * 4;
    }
    p.input_elems = off;
    const size_t smem = 16 + ((size_t)ipc * p.inst_stride + off + F.total) * sizeof(T);
    if (smem > 227 * 1024) return QPMPC_B200_ESHAPE;
    launch((p.batch + ipc - 1) / ipc, wpc * 32, smem, [&]() { mpc_solve_kernel<T, NP, 1, true, false, true, true>(p); });
    return 0;
}

// factor_np / factor_has_ft of qpmpc_b200.cu
int emu_factor_np(const qpmpc_b200_desc *d) {
    auto shared = [](int mode, bool opt) {
        return mode == QPMPC_B200_SHARED_LTI || mode == QPMPC_B200_SHARED_LTV || (opt && mode == QPMPC_B200_ABSENT);
    };
    if (!d || d->nc <= 0 || (d->nc & 1) || !d->paired || d->method != QPMPC_B200_ACTIVE_SET) return 0;
    if (!shared(d->mode_A, false) || !shared(d->mode_B, false) || !shared(d->mode_C, true) || !shared(d->mode_D, true))
        return 0;
    Variant v;
    if (!pick_variant(d->N * d->nu, d->N * d->nc, &v, true) || !v.paired) return 0;
    return v.np;
}

// the active-set kernels of one dtype: warp kernel variants, or the CTA kernel
template <typename T>
int emu_solve_t(const qpmpc_b200_desc *d, SolveParams &p, int wpc) {
    Variant v;
    // as solve_impl / rows_paired (qpmpc_b200.cu)
    const bool paired = d->paired != 0 && d->nc > 0 && (d->nc & 1) == 0 && env_int("QPMPC_B200_NO_PAIRED", 0) == 0;
    if constexpr (sizeof(T) == 8) {
        const int lr = env_int("QPMPC_B200_LR", -1);
        if (lr != 0 && env_int("QPMPC_B200_FORCE_CTA", 0) == 0 && lr_applicable(p, paired) &&
            (p.n > 16 || (p.n > 8 && env_int("QPMPC_B200_LR16", LR16_DEFAULT) != 0)))
            return p.n <= 16 ? solve_lr<T, 16>(p) : p.n <= 32 ? solve_lr<T, 32>(p) : solve_lr<T, 64>(p);
    }
    const bool warp_ok = pick_variant(p.n, p.m, &v, paired);
    if (!warp_ok || env_int("QPMPC_B200_FORCE_CTA", 0) != 0) {
        int threads = env_int("QPMPC_B200_CTA_THREADS", 256);
        threads = threads < 32 ? 32 : (threads > 256 ? 256 : (threads & ~31));
        return solve_cta<T>(p, threads);
    }
    if (v.paired) {
        if (wpc <= 0) wpc = v.np == 32 ? 4 : 8;
        if (v.np == 8) return solve_variant<T, 8, 1, true, false, true>(p, wpc);
        if (v.np == 16) return solve_variant<T, 16, 1, true, false, true>(p, wpc);
        return solve_variant<T, 32, 1, true, false, true>(p, wpc);
    }
    if (wpc <= 0) wpc = 8;
    switch (v.np * 10 + v.mr) {
        case 82: return solve<T, 8, 2, true>(p, wpc);
        case 84: return solve<T, 8, 4, true>(p, wpc);
        case 162: return solve<T, 16, 2, true>(p, wpc);
        case 164: return solve<T, 16, 4, false>(p, wpc);
        case 322: return solve<T, 32, 2, false>(p, wpc);
        case 324: return solve<T, 32, 4, false>(p, wpc);
    }
    return QPMPC_B200_ESHAPE;
}

}  // namespace

extern "C" {

int pdip_emu_solve(int dtype, int np, int mr, int count, int n, int m, const double *P, const double *q,
                   const double *G, const double *h, int max_iter, double tol, int polish, double pmin, double *U,
                   double *Z, int *status, int *iters) {
    // as launch_pdip; the float instantiation (not offered through the ABI) only works with the
    // substitutions: an explicit L^-1 of the late, ill-conditioned H is beyond single precision
    const bool ls = dtype == 0 && env_int("QPMPC_B200_PDIP_SOLVE", 0) != 0;
#define CASE(T, NP, MR)                                                                                             \
    if (np == NP && mr == MR)                                                                                       \
        return ls ? pdip_core_run<T, NP, MR, true>(count, n, m, P, q, G, h, max_iter, tol, polish, pmin, U, Z,     \
                                                   status, iters)                                                   \
                  : pdip_core_run<T, NP, MR, false>(count, n, m, P, q, G, h, max_iter, tol, polish, pmin, U, Z,    \
                                                    status, iters);
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

// Self-test of the emulator's convergence check: a group-wise reduction entered
// by all lanes (which = 0), or hidden behind a short-circuited `&&` whose left
// side differs between the two 16-lane groups of the warp (which = 1: the bug
// class that hangs a device).  Returns the number of divergent collectives seen.
long emu_selftest(int which) {
    const long before = emu::divergent_collectives();
    emu::run_cta(32, emu::Idx{0, 0, 0}, emu::Idx{1, 1, 1}, [&]() {
        const int lane = emu::lane_id();
        const bool flag = lane < 16;  // uniform inside a group, different across groups
        const double v = (double)lane;
        bool ok;
        if (which == 0) {
            const double r = pdip_max<double, 16>(v);
            ok = flag && r > 3.0;
        } else {
            ok = flag && pdip_max<double, 16>(v) > 3.0;
        }
        __syncwarp();
        (void)ok;
    });
    return emu::divergent_collectives() - before;
}

int emu_smem_overruns(void) { return g_smem_overrun; }
void emu_set_lane_order(int descending) { emu::lane_order() = descending; }
long emu_divergent_collectives(void) { return emu::divergent_collectives(); }
long emu_sync_points(void) { return emu::sync_points(); }
// per-source-line histogram of synchronisation points (line numbers of the kernel headers; the
// files overlap in line numbers, the caller knows which kernel it ran); reset = 1 clears it
long emu_sync_points_at(int line, int reset) {
    auto &h = emu::sync_points_by_line();
    if (reset) std::fill(h.begin(), h.end(), 0L);
    return (line >= 0 && line < (int)h.size()) ? h[line] : 0;
}

// qpmpc_b200_solve with host pointers.  `wpc`: warps per CTA (<= 0: the launch default).
int emu_solve(const qpmpc_b200_desc *d, const qpmpc_b200_operands *in, const qpmpc_b200_outputs *out, int wpc) {
    int rc = check_desc(d, in);
    if (rc) return rc;
    if (!out || !out->U || !out->status) return QPMPC_B200_EINVAL;
    if (d->batch == 0) return 0;
    SolveParams p;
    fill_params(d, in, &p);
    p.U = out->U;
    p.status = out->status;
    p.iters = out->iters;
    p.Z = out->Z;
    if (d->method == QPMPC_B200_PDIP) {
        // as solve_impl: double precision, warp kernel shapes only
        Variant v;
        if (d->dtype != QPMPC_B200_F64 || !pick_variant(p.n, p.m, &v)) return QPMPC_B200_EUNSUPPORTED;
        p.max_iter = d->max_iter > 0 ? d->max_iter : 50;
        const int polish = (d->flags & QPMPC_B200_FLAG_NO_POLISH) ? 0 : 1;
        if (wpc <= 0) wpc = 4;
        switch (v.np * 10 + v.mr) {
            case 82: return pdip<double, 8, 2>(p, polish, wpc);
            case 84: return pdip<double, 8, 4>(p, polish, wpc);
            case 162: return pdip<double, 16, 2>(p, polish, wpc);
            case 164: return pdip<double, 16, 4>(p, polish, wpc);
            case 322: return pdip<double, 32, 2>(p, polish, wpc);
            case 324: return pdip<double, 32, 4>(p, polish, wpc);
        }
        return QPMPC_B200_ESHAPE;
    }
    return d->dtype == QPMPC_B200_F64 ? emu_solve_t<double>(d, p, wpc) : emu_solve_t<float>(d, p, wpc);
}

// qpmpc_b200_condense with host pointers.
int emu_condense(const qpmpc_b200_desc *d, const qpmpc_b200_operands *in, const qpmpc_b200_qp_fields *out) {
    int rc = check_desc(d, in);
    if (rc) return rc;
    if (!out || d->dtype != QPMPC_B200_F64) return QPMPC_B200_EINVAL;
    if (d->batch == 0) return 0;
    SolveParams p;
    fill_params(d, in, &p);
    p.P = out->P;
    p.q = out->q;
    p.G = out->G;
    p.h = out->h;
    p.Phi = out->Phi;
    p.Psi = out->Psi;
    p.phi_last = out->phi_last;
    p.psi_last = out->psi_last;
    Variant v;
    if (!pick_variant(p.n, p.m, &v) || env_int("QPMPC_B200_FORCE_CTA", 0) != 0) return condense_cta<double>(p);
    switch (v.np * 10 + v.mr) {
        case 82: return condense<double, 8, 2>(p, 2);
        case 84: return condense<double, 8, 4>(p, 2);
        case 162: return condense<double, 16, 2>(p, 2);
        case 164: return condense<double, 16, 4>(p, 2);
        case 322: return condense<double, 32, 2>(p, 2);
        case 324: return condense<double, 32, 4>(p, 2);
    }
    return QPMPC_B200_ESHAPE;
}

// qpmpc_b200_integrate with host pointers (mpc_integrate_kernel).
int emu_integrate(const qpmpc_b200_desc *d, const qpmpc_b200_operands *in, const void *U, void *X) {
    if (!d || !in || !U || !X || !in->A || !in->B || !in->x0 || d->dtype != QPMPC_B200_F64) return QPMPC_B200_EINVAL;
    IntegrateParams p;
    p.batch = d->batch;
    p.N = d->N;
    p.nx = d->nx;
    p.nu = d->nu;
    p.A = in->A;
    p.B = in->B;
    p.x0 = in->x0;
    p.U = U;
    p.X = X;
    auto ltv = [](int mode) { return mode == QPMPC_B200_SHARED_LTV || mode == QPMPC_B200_BATCH_LTV; };
    auto per = [](int mode) { return mode == QPMPC_B200_BATCH_LTI || mode == QPMPC_B200_BATCH_LTV; };
    p.sA = ltv(d->mode_A) ? d->nx * d->nx : 0;
    p.sB = ltv(d->mode_B) ? d->nx * d->nu : 0;
    p.bA = per(d->mode_A) ? (long long)d->nx * d->nx * (ltv(d->mode_A) ? d->N : 1) : 0;
    p.bB = per(d->mode_B) ? (long long)d->nx * d->nu * (ltv(d->mode_B) ? d->N : 1) : 0;
    p.bx0 = d->mode_x0 == QPMPC_B200_VEC_BATCH ? d->nx : 0;
    const int threads = 128, grid = (d->batch + threads - 1) / threads;
    launch(grid, threads, 0, [&]() { mpc_integrate_kernel<double>(p); });
    return 0;
}

// qpmpc_b200_pendulum_closed_loop with host pointers: pendulum_step_kernel and the
// fused solve alternate exactly as in qpmpc_b200.cu.
// qpmpc_b200_factor_bytes / _factor / _solve_factored with host pointers (double only)
size_t emu_factor_bytes(const qpmpc_b200_desc *d) {
    const int np = emu_factor_np(d);
    if (!np || d->dtype != QPMPC_B200_F64) return 0;
    SolveParams sp;
    qpmpc_b200_operands none = {};
    fill_params(d, &none, &sp);
    const size_t n = (size_t)d->N * d->nu, m = (size_t)d->N * d->nc, nx = d->nx, N = d->N;
    return ((size_t)factor_layout(np, d->nx, d->N, sp.q_wx != 0).total + n * n + m * n + N * nx * nx + N * nx * n +
            nx * nx + nx * n + 16) * 8;
}
int emu_factor(const qpmpc_b200_desc *d, const qpmpc_b200_operands *in, void *record) {
    int rc = check_desc(d, in);
    if (rc) return rc;
    const int np = emu_factor_np(d);
    if (!np || d->dtype != QPMPC_B200_F64 || !record) return QPMPC_B200_EUNSUPPORTED;
    SolveParams sp;
    fill_params(d, in, &sp);
    const FactorLay F = factor_layout(np, d->nx, d->N, sp.q_wx != 0);
    const size_t n = (size_t)d->N * d->nu, m = (size_t)d->N * d->nc, nx = d->nx, N = d->N;
    double *scr = static_cast<double *>(record) + F.total;
    qpmpc_b200_qp_fields f;
    f.P = scr, scr += n * n;
    f.G = scr, scr += m * n;
    f.Phi = scr, scr += N * nx * nx;
    f.Psi = scr, scr += N * nx * n;
    f.phi_last = scr, scr += nx * nx;
    f.psi_last = scr;
    f.q = f.h = nullptr;
    qpmpc_b200_desc d1 = *d;
    d1.batch = 1;
    if ((rc = emu_condense(&d1, in, &f)) != 0) return rc;
    FactorParams fp;
    fp.N = d->N, fp.nx = d->nx, fp.nu = d->nu, fp.nc = d->nc, fp.n = (int)n, fp.m = (int)m, fp.NP = np;
    fp.has_ft = sp.q_wx, fp.q_wt = sp.q_wt, fp.q_wx = sp.q_wx;
    fp.w_t = sp.w_t, fp.w_x = sp.w_x;
    fp.P = f.P, fp.G = f.G, fp.Phi = f.Phi, fp.Psi = f.Psi, fp.phi_last = f.phi_last, fp.psi_last = f.psi_last;
    fp.C = d->mode_C == QPMPC_B200_ABSENT ? nullptr : in->C;
    fp.stepC = sp.op[OP_C].step;
    fp.record = record;
    launch(1, 128, FACTOR_SMEM_BYTES, [&]() { mpc_factor_kernel<double>(fp); });
    return 0;
}
static int emu_solve_factored_impl(const qpmpc_b200_desc *d, const qpmpc_b200_operands *in, const void *record,
                                   const qpmpc_b200_outputs *out, const LoopDev *loop, int wpc) {
    int rc = check_desc(d, in);
    if (rc) return rc;
    const int np = emu_factor_np(d);
    if (!np || d->dtype != QPMPC_B200_F64 || !record || !out || !out->U || !out->status) return QPMPC_B200_EUNSUPPORTED;
    if (d->batch == 0) return 0;
    SolveParams p;
    fill_params(d, in, &p);
    p.U = out->U;
    p.status = out->status;
    p.iters = out->iters;
    p.Z = out->Z;
    {   // (qpmpc_b200.cu: solve_factored_impl -- long terminal-cost horizons take the structure-exploiting kernel)
        const int lr = env_int("QPMPC_B200_LR", -1);
        const bool paired = d->paired != 0 && d->nc > 0 && (d->nc & 1) == 0 && env_int("QPMPC_B200_NO_PAIRED", 0) == 0;
        if (!loop && d->method == QPMPC_B200_ACTIVE_SET && lr != 0 && env_int("QPMPC_B200_FORCE_CTA", 0) == 0 &&
            lr_applicable(p, paired) && (p.n > 16 || (p.n > 8 && env_int("QPMPC_B200_LR16", LR16_DEFAULT) != 0)) &&
            env_int("QPMPC_B200_FACTORED_LR", 1) != 0)
            return emu_solve(d, in, out, wpc);
    }
    p.record = record;
    if (loop) p.loop = *loop;
    if (np == 8) return solve_pre<double, 8>(p, wpc);
    if (np == 16) return solve_pre<double, 16>(p, wpc);
    return solve_pre<double, 32>(p, wpc);
}
int emu_solve_factored(const qpmpc_b200_desc *d, const qpmpc_b200_operands *in, const void *record,
                       const qpmpc_b200_outputs *out, int wpc) {
    return emu_solve_factored_impl(d, in, record, out, nullptr, wpc);
}
// (qpmpc_b200.cu: loop_fused)
static bool emu_loop_fused(const void *record, int cycles) {
    return record && cycles > 0 && env_int("QPMPC_B200_LOOP_FUSED", 1) != 0;
}

int emu_pendulum_closed_loop(const qpmpc_b200_desc *d, const qpmpc_b200_operands *in, const qpmpc_b200_outputs *out,
                             const qpmpc_b200_closed_loop *loop) {
    if (!d || !in || !out || !loop || d->nx != 4 || d->nu != 1 || d->dtype != QPMPC_B200_F64) return QPMPC_B200_EINVAL;
    PendulumStepParams pp;
    pp.batch = d->batch;
    pp.N = d->N;
    pp.n = d->N * d->nu;
    pp.dt = loop->dt;
    pp.T = loop->sampling_period;
    pp.g = loop->gravity;
    pp.omega2 = loop->gravity / loop->length;
    pp.state = const_cast<void *>(in->x0);
    pp.U = out->U;
    pp.status = out->status;
    pp.v_target = loop->v_target;
    pp.goal = const_cast<void *>(in->goal);
    pp.targets = const_cast<void *>(in->targets);
    pp.unsolved = loop->unsolved;
    pp.upright = loop->upright;
    pp.iters = out->iters;
    pp.iter_sum = nullptr;
    const int threads = 128, grid = (d->batch + threads - 1) / threads;
    auto step = [&](int substeps, int slot) {
        pp.substeps = substeps;
        pp.iter_sum = (loop->iterations && out->iters && slot > 0) ? (long long *)loop->iterations + (slot - 1) : nullptr;
        pp.traj = loop->trajectory ? static_cast<char *>(loop->trajectory) + (size_t)slot * d->batch * 4 * 8 : nullptr;
        launch(grid, threads, 0, [&]() { pendulum_step_kernel<double>(pp); });
    };
    step(0, 0);
    if (emu_loop_fused(loop->record, loop->cycles)) {
        LoopDev ld;
        std::memset(&ld, 0, sizeof(ld));
        ld.kind = 1, ld.cycles = loop->cycles, ld.substeps = loop->substeps;
        ld.dt = pp.dt, ld.T = pp.T, ld.omega2 = pp.omega2, ld.g = pp.g;
        ld.state = pp.state, ld.v_target = pp.v_target, ld.goal = pp.goal, ld.targets = pp.targets;
        ld.traj = loop->trajectory;
        ld.unsolved = loop->unsolved, ld.upright = loop->upright;
        ld.iter_sum = (loop->iterations && out->iters) ? (long long *)loop->iterations : nullptr;
        return emu_solve_factored_impl(d, in, loop->record, out, &ld, 0);
    }
    for (int c = 0; c < loop->cycles; ++c) {
        const int rc = loop->record ? emu_solve_factored(d, in, loop->record, out, 0) : emu_solve(d, in, out, 0);
        if (rc) return rc;
        step(loop->substeps, c + 1);
    }
    return 0;
}

int emu_lipm_closed_loop(const qpmpc_b200_desc *d, const qpmpc_b200_operands *in, const qpmpc_b200_outputs *out,
                         const qpmpc_b200_lipm_loop *loop) {
    if (!d || !in || !out || !loop || d->nx != 3 || d->nu != 1 || d->nc != 2 || d->dtype != QPMPC_B200_F64)
        return QPMPC_B200_EINVAL;
    LipmStepParams pp;
    pp.batch = d->batch, pp.N = d->N, pp.n = d->N * d->nu;
    pp.nb_dsp = loop->nb_dsp_steps, pp.nb_ssp = loop->nb_ssp_steps;
    pp.dt = loop->sampling_period / loop->substeps;
    pp.foot_size = loop->foot_size, pp.max_zmp = loop->max_zmp_dist;
    pp.state = const_cast<void *>(in->x0);
    pp.U = out->U, pp.status = out->status;
    pp.support_foot = loop->support_foot, pp.strides = loop->strides;
    pp.phase_index = loop->phase_index, pp.stride_index = loop->stride_index;
    pp.e = const_cast<void *>(in->e), pp.goal = const_cast<void *>(in->goal);
    pp.unsolved = loop->unsolved;
    const int threads = 128, grid = (d->batch + threads - 1) / threads;
    auto step = [&](int substeps, int slot) {
        pp.substeps = substeps;
        pp.traj = loop->trajectory ? static_cast<char *>(loop->trajectory) + (size_t)slot * d->batch * 3 * 8 : nullptr;
        launch(grid, threads, 0, [&]() { lipm_step_kernel<double>(pp); });
    };
    step(0, 0);
    if (emu_loop_fused(loop->record, loop->cycles)) {
        LoopDev ld;
        std::memset(&ld, 0, sizeof(ld));
        ld.kind = 2, ld.cycles = loop->cycles, ld.substeps = loop->substeps;
        ld.dt = pp.dt, ld.nb_dsp = pp.nb_dsp, ld.nb_ssp = pp.nb_ssp;
        ld.foot_size = pp.foot_size, ld.max_zmp = pp.max_zmp;
        ld.state = pp.state, ld.goal = pp.goal, ld.e = pp.e;
        ld.support_foot = pp.support_foot, ld.strides = pp.strides;
        ld.phase_index = pp.phase_index, ld.stride_index = pp.stride_index;
        ld.traj = loop->trajectory;
        ld.unsolved = loop->unsolved;
        return emu_solve_factored_impl(d, in, loop->record, out, &ld, 0);
    }
    for (int c = 0; c < loop->cycles; ++c) {
        const int rc = loop->record ? emu_solve_factored(d, in, loop->record, out, 0) : emu_solve(d, in, out, 0);
        if (rc) return rc;
        step(loop->substeps, c + 1);
    }
    return 0;
}

}  // extern "C"
