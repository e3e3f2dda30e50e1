// mpc_launch.cuh -- launch wrappers of the warp kernels (mpc_kernels.cuh).
//
// The kernels are instantiated in one small translation unit per (dtype, NP)
// (inst_*.cu) so that the library builds in parallel; the host API
// (qpmpc_b200.cu) only sees the declarations below.  A translation unit that
// defines QPMPC_INSTANTIATE gets the definitions.
#pragma once

#include "../../include/qpmpc_b200.h"
#include "mpc_common.cuh"

namespace qpmpc {

void count_launch();                        // qpmpc_b200.cu

template <typename T, int NP, int MR, bool MREG>
int launch_solve(SolveParams p, cudaStream_t stream);
template <typename T, int NP>  // paired rows (desc.paired): one stored row per lane stands for [G+; -G+]
int launch_solve_paired(SolveParams p, cudaStream_t stream);
template <typename T, int NP, int MR>
int launch_condense(SolveParams p, cudaStream_t stream);
template <typename T, int NP>  // structure-exploiting kernel for terminal-cost problems (mpc_lr_kernel.cuh)
int launch_solve_lr(SolveParams p, cudaStream_t stream);
template <typename T, int NP>  // shared-model fast path (mpc_factor.cuh): p.record holds the model's factor
int launch_solve_pre(SolveParams p, cudaStream_t stream);
template <typename T, int NP, int MR>
int launch_pdip(SolveParams p, int polish, cudaStream_t stream);  // mpc_pdip.cuh

}  // namespace qpmpc

#ifdef QPMPC_INSTANTIATE
#include "mpc_kernels.cuh"
#include "mpc_pdip.cuh"

namespace qpmpc {

template <typename T, int NP, int MR, bool MREG, bool RS, bool PAIRED = false>
int launch_solve_variant(SolveParams p, cudaStream_t stream) {
    using L = Lay<T, NP, MR, MREG, RS, PAIRED>;
    constexpr int IPW = 32 / NP;
    int wpc = env_int("QPMPC_B200_WPC", (PAIRED && NP == 32) ? 4 : 8);
    if (wpc < 1) wpc = 1;
    if (wpc > 8) wpc = 8;
    size_t smem = 0;
    for (;; --wpc) {
        smem = layout_smem<T>(&p, L::fixed, L::szG, NP, IPW * wpc, MREG);
        if (smem <= 227 * 1024 || wpc == 1) break;
    }
    if (smem > 227 * 1024) return QPMPC_B200_ESHAPE;
    // Occupancy experiments only: pad the dynamic shared memory request.
    const size_t pad = (size_t)env_int("QPMPC_B200_SMEM_PAD_KB", 0) * 1024;
    if (pad && smem + pad <= 227 * 1024) smem += pad;
    const int ipc = IPW * wpc;
    auto kern = mpc_solve_kernel<T, NP, MR, MREG, RS, PAIRED>;
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return (int)err;
    const int grid = (p.batch + ipc - 1) / ipc;
    if (grid == 0) return 0;
    kern<<<grid, wpc * 32, smem, stream>>>(p);
    count_launch();
    return (int)cudaGetLastError();
}

// The NP = 16 register-resident kernel exists in two builds: with the per-row
// constants of the iteration in shared memory (8 fewer live registers: +4 % on
// iteration-dominated workloads such as config 2) or in registers (smaller
// shared-memory footprint, hence a larger L1 for the spills of the heavier
// condensing phase: +10 % on stage-cost problems such as config 3).
template <typename T, int NP, int MR, bool MREG>
int launch_solve(SolveParams p, cudaStream_t stream) {
    if constexpr (NP == 16 && MREG) {
        const int rs = env_int("QPMPC_B200_ROWS_SMEM", -1);
        if (rs > 0 || (rs < 0 && !p.has_wx)) return launch_solve_variant<T, NP, MR, MREG, true>(p, stream);
    }
    return launch_solve_variant<T, NP, MR, MREG, false>(p, stream);
}

template <typename T, int NP>
int launch_solve_paired(SolveParams p, cudaStream_t stream) {
    return launch_solve_variant<T, NP, 1, true, false, true>(p, stream);
}

// Shared-model fast path: per-instance region = Lay<PRE>::fixed, then the CTA's inputs (only the
// operands that still vary per solve: e, x0, goal, targets), then the record of the model.
template <typename T, int NP>
int launch_solve_pre(SolveParams p, cudaStream_t stream) {
    using L = Lay<T, NP, 1, true, false, true, true>;
    constexpr int IPW = 32 / NP;
    for (int o : {OP_A, OP_B, OP_C, OP_D}) p.op[o] = OperandView{nullptr, 0, 0, 0, 0};  // in the record
    p.toeplitz = 0;
    p.inst_stride = (L::fixed + 3) / 4 * 4;
    p.gt_off = p.g_off = p.scr_off = 0;
    const FactorLay F = factor_layout(NP, p.nx, p.N, p.q_wx != 0);
    // (a fused loop copies the record once per CTA for all its cycles: smaller CTAs cost nothing and
    // even out the last wave -- measured 382 -> 399 M solves/s on config 3)
    int wpc = env_int("QPMPC_B200_WPC", p.loop.kind != 0 ? 4 : 8);
    wpc = wpc < 1 ? 1 : (wpc > 8 ? 8 : wpc);
    size_t smem = 0;
    for (;; --wpc) {
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
        smem = 16 + ((size_t)ipc * p.inst_stride + off + F.total) * sizeof(T);
        if (smem <= 227 * 1024 || wpc == 1) break;
    }
    if (smem > 227 * 1024) return QPMPC_B200_ESHAPE;
    const int ipc = IPW * wpc;
    auto kern = mpc_solve_kernel<T, NP, 1, true, false, true, true>;
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return (int)err;
    const int grid = (p.batch + ipc - 1) / ipc;
    if (grid == 0) return 0;
    kern<<<grid, wpc * 32, smem, stream>>>(p);
    count_launch();
    return (int)cudaGetLastError();
}

template <typename T, int NP, int MR>
int launch_condense(SolveParams p, cudaStream_t stream) {
    using L = Lay<T, NP, MR, false>;
    constexpr int IPW = 32 / NP;
    int wpc = 2;
    size_t smem = 0;
    for (;; --wpc) {
        smem = layout_smem<T>(&p, L::fixed, L::szG, NP, IPW * wpc, false);
        if (smem <= 227 * 1024 || wpc == 1) break;
    }
    if (smem > 227 * 1024) return QPMPC_B200_ESHAPE;
    const int ipc = IPW * wpc;
    const int grid = (p.batch + ipc - 1) / ipc;
    if (grid == 0) return 0;
    // P, q, G, h come from the same phase-A code the fused kernel runs; the
    // Phi / Psi stacks need the instrumented variant: two launches if both.
    const bool want_qp = p.P || p.q || p.G || p.h;
    const bool want_dump = p.Phi || p.Psi || p.phi_last || p.psi_last;
    if (want_qp) {
        SolveParams a = p;
        a.Phi = a.Psi = a.phi_last = a.psi_last = nullptr;
        auto kern = mpc_condense_kernel<T, NP, MR, false>;
        cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (err != cudaSuccess) return (int)err;
        kern<<<grid, wpc * 32, smem, stream>>>(a);
        count_launch();
    }
    if (want_dump) {
        SolveParams b = p;
        b.P = b.q = b.G = b.h = nullptr;
        auto kern = mpc_condense_kernel<T, NP, MR, true>;
        cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (err != cudaSuccess) return (int)err;
        kern<<<grid, wpc * 32, smem, stream>>>(b);
        count_launch();
    }
    return (int)cudaGetLastError();
}

// Interior-point kernel (desc.method = QPMPC_B200_PDIP): dense G, one work
// region of PdipLay::fixed + szG (+ generic-condensing scratch) per instance.
template <typename T, int NP, int MR>
int launch_pdip(SolveParams p, int polish, cudaStream_t stream) {
    using L = PdipLay<T, NP, MR>;
    constexpr int IPW = 32 / NP;
    int wpc = env_int("QPMPC_B200_PDIP_WPC", 4);
    if (wpc < 1) wpc = 1;
    if (wpc > 8) wpc = 8;
    size_t smem = 0;
    for (;; --wpc) {
        smem = layout_smem<T>(&p, L::fixed, L::szG, NP, IPW * wpc, false, true);
        if (smem <= 227 * 1024 || wpc == 1) break;
    }
    if (smem > 227 * 1024) return QPMPC_B200_ESHAPE;
    const int ipc = IPW * wpc;
    // The two substitutions per solve: by 2 NP dependent shuffles (default: backward stable), or
    // through L^-1 and shared memory (LS, QPMPC_B200_PDIP_SOLVE=1: 4 __syncwarp per solve, but
    // the explicit inverse of the late, ill-conditioned H costs robustness at tight tolerances).
    const bool ls = env_int("QPMPC_B200_PDIP_SOLVE", 0) != 0;
    auto kern = ls ? mpc_pdip_kernel<T, NP, MR, true> : mpc_pdip_kernel<T, NP, MR, false>;
    cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return (int)err;
    const int grid = (p.batch + ipc - 1) / ipc;
    if (grid == 0) return 0;
    kern<<<grid, wpc * 32, smem, stream>>>(p, polish);
    count_launch();
    return (int)cudaGetLastError();
}

#define QPMPC_INSTANTIATE_VARIANT(T, NP, MR, MREG)                               \
    template int launch_solve<T, NP, MR, MREG>(SolveParams, cudaStream_t);      \
    template int launch_condense<T, NP, MR>(SolveParams, cudaStream_t);
#define QPMPC_INSTANTIATE_PAIRED(T, NP)                                         \
    template int launch_solve_paired<T, NP>(SolveParams, cudaStream_t);         \
    template int launch_solve_pre<T, NP>(SolveParams, cudaStream_t);
// the interior-point kernel is double precision only (qpmpc_b200.cu:solve_impl)
#define QPMPC_INSTANTIATE_PDIP(NP, MR) template int launch_pdip<double, NP, MR>(SolveParams, int, cudaStream_t);

}  // namespace qpmpc
#endif  // QPMPC_INSTANTIATE
