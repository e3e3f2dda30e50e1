// mpc_common.cuh -- parameter blocks, shared-memory layout and small device
// helpers shared by the kernels of the batched MPC engine (sm_100a only).
#pragma once

// QPMPC_HOST_EMU: the translation unit is a host build for the fiber warp
// emulator (tests/emu/): the CUDA surface comes from tests/emu/warp_emu.h,
// included first, and the inline-PTX helpers below get host bodies.
#ifndef QPMPC_HOST_EMU
#include <cuda_runtime.h>
#endif
#include <stdint.h>

namespace qpmpc {

constexpr unsigned FULL_MASK = 0xffffffffu;

int env_int(const char *name, int dflt);  // integer environment knob (qpmpc_b200.cu)

// Operand slots in the staged-input table.
enum { OP_A = 0, OP_B, OP_C, OP_D, OP_E, OP_X0, OP_GOAL, OP_TGT, OP_COUNT };

// One staged operand: where it lives in HBM and where a CTA puts it in shared
// memory.  `sz` elements per instance (item * N for per-step operands).
struct OperandView {
    const void *ptr;   // nullptr: absent
    int sz;            // elements staged per instance (or once if shared)
    int step;          // element stride between steps in the staged copy (0: LTI)
    int per_instance;  // 1: [batch, sz] in HBM; 0: one copy shared by the batch
    int smem_off;      // element offset inside the CTA's input region
};

struct SolveParams {
    int batch, N, nx, nu, nc, n, m;
    OperandView op[OP_COUNT];
    int has_wt, has_wx;  // weight "is not None": term enters P (mpc_qp.py:102,104)
    int q_wt, q_wx;      // term enters q (weight > 1e-10 and reference present)
    double w_t, w_x, w_u;
    int max_iter;
    double tol;
    // shared-memory geometry (elements of T), computed by the host
    int inst_stride;     // per-instance work region
    int toeplitz;        // 1: A, B, C time-invariant and nx in registers -> G kept as a table
    int gt_off, g_off, scr_off;  // tail regions of the work region (TailLay)
    int input_elems;     // CTA-level input region
    int present_mask;    // bit o set: operand o is present (staged)
    // outputs
    void *U;
    int *status;
    int *iters;
    void *Z;
    // peer destinations of the fused gather (qpmpc_b200_solve_scatter): rows
    // [row_off, row_off + batch) of every peer buffer receive U and status
    int npeers;
    long long row_off;
    void *peerU[8];
    int *peer_status[8];
    // condensed-field dump (condense kernel only)
    void *P, *q, *G, *h, *Phi, *Psi, *phi_last, *psi_last;
};

template <typename T> struct Pair;
template <> struct Pair<double> { using type = double2; };
template <> struct Pair<float> { using type = float2; };

__device__ __forceinline__ double rsqrt_(double v) { return rsqrt(v); }
__device__ __forceinline__ float rsqrt_(float v) { return rsqrtf(v); }
__device__ __forceinline__ double sqrt_(double v) { return sqrt(v); }
__device__ __forceinline__ float sqrt_(float v) { return sqrtf(v); }
// Branch-free reciprocal and reciprocal square root for well-scaled positive
// (rcp: non-zero) arguments: the SFU seed (about 20 bits) plus two Newton
// steps, ~1 ulp.  The library versions spend most of their instructions on
// subnormal / infinity handling the solver never needs; zero, subnormal and
// negative arguments give inf / NaN, which the callers test for.
__device__ __forceinline__ double rcp_(double v) {
#ifdef QPMPC_HOST_EMU
    return 1.0 / v;
#else
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(v));
    double e = fma(-v, r, 1.0);
    r = fma(r, e, r);
    e = fma(-v, r, 1.0);
    return fma(r, e, r);
#endif
}
__device__ __forceinline__ float rcp_(float v) { return __frcp_rn(v); }
__device__ __forceinline__ double frsqrt_(double v) {
#ifdef QPMPC_HOST_EMU
    return 1.0 / sqrt(v);
#else
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(v));
    const double h = 0.5 * v;
    double e = fma(-h * y, y, 0.5);
    y = fma(y, e, y);
    e = fma(-h * y, y, 0.5);
    return fma(y, e, y);
#endif
}
__device__ __forceinline__ float frsqrt_(float v) { return rsqrtf(v); }

// Exact minimum / maximum of non-negative values over the lanes in `mask`
// (every lane of the mask calls with the same mask): one or two REDUX.
__device__ __forceinline__ double seg_min_pos(double v, unsigned mask) {
    const unsigned hi = (unsigned)__double2hiint(v), lo = (unsigned)__double2loint(v);
    const unsigned mh = __reduce_min_sync(mask, hi);
    const unsigned ml = __reduce_min_sync(mask, hi == mh ? lo : 0xffffffffu);
    return __hiloint2double((int)mh, (int)ml);
}
__device__ __forceinline__ float seg_min_pos(float v, unsigned mask) {
    return __uint_as_float(__reduce_min_sync(mask, __float_as_uint(v)));
}
__device__ __forceinline__ double seg_max_pos(double v, unsigned mask) {
    const unsigned hi = (unsigned)__double2hiint(v), lo = (unsigned)__double2loint(v);
    const unsigned mh = __reduce_max_sync(mask, hi);
    const unsigned ml = __reduce_max_sync(mask, hi == mh ? lo : 0u);
    return __hiloint2double((int)mh, (int)ml);
}
__device__ __forceinline__ float seg_max_pos(float v, unsigned mask) {
    return __uint_as_float(__reduce_max_sync(mask, __float_as_uint(v)));
}
__device__ __forceinline__ double abs_(double v) { return fabs(v); }
__device__ __forceinline__ float abs_(float v) { return fabsf(v); }

template <typename T> struct Num;
template <> struct Num<double> {
    static __device__ __forceinline__ double inf() { return __longlong_as_double(0x7ff0000000000000LL); }
    static __device__ __forceinline__ double nan() { return __longlong_as_double(0x7ff8000000000000LL); }
    static constexpr double viol_eps = 1e-13;  // "is this row violated" (oracle/mpc_oracle.c)
    static constexpr double dep_eps = 1e-24;   // |d2|^2 <= dep_eps |d|^2: dependent normal
};
template <> struct Num<float> {
    static __device__ __forceinline__ float inf() { return __int_as_float(0x7f800000); }
    static __device__ __forceinline__ float nan() { return __int_as_float(0x7fc00000); }
    static constexpr float viol_eps = 2e-6f;
    static constexpr float dep_eps = 1e-10f;
};

#ifdef QPMPC_HOST_EMU
// ---- host stand-ins: the copy happens at issue time, the barrier is a no-op --
__device__ __forceinline__ void mbar_init(uint64_t *, unsigned) {}
__device__ __forceinline__ void fence_barrier_init() {}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *, unsigned) {}
__device__ __forceinline__ void mbar_wait(uint64_t *, unsigned) {}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, uint64_t *) {
    memcpy(dst, src, bytes);
}
#else
// ---- mbarrier + 1-D bulk TMA (cp.async.bulk) -------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, unsigned parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
// global -> shared bulk copy, completion reported to an mbarrier (SASS: UBLKCP).
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

#endif  // QPMPC_HOST_EMU

}  // namespace qpmpc
