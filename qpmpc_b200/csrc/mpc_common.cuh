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

// A whole receding-horizon loop inside ONE launch of the shared-model solve kernel
// (mpc_kernels.cuh, PRE): the instances are independent, so each lane group runs all its cycles
// -- solve, apply the first input to the plant, write the next cycle's vectors into the staged
// inputs in shared memory -- without leaving the SM (mpc_plant.cuh has the plant pieces).
struct LoopDev {
    int kind;                         // 0: a single solve; 1: wheeled inverted pendulum; 2: LIPM walking
    int cycles, substeps;
    double dt, T, omega2, g;          // plant step; pendulum: MPC sampling period, g / length, g
    int nb_dsp, nb_ssp;               // walking: steps of a double / single support phase
    double foot_size, max_zmp;
    void *state;                      // [batch, nx] in/out
    const void *v_target;             // pendulum [batch]
    void *support_foot;               // walking [batch] in/out
    const void *strides;              // walking [batch, 2]
    int *phase_index, *stride_index;  // walking [batch] in/out
    void *goal, *targets, *e;         // vectors of the NEXT cycle, written back after the last one
    void *traj;                       // optional [cycles + 1, batch, nx] (slot 0 is the caller's)
    int *unsolved, *upright;          // optional counters
    long long *iter_sum;              // optional [cycles]
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
    // shared-model fast path (mpc_factor.cuh): the record of the model, staged once per CTA
    const void *record;
    // condense-only CTA kernel: 1 = do not accumulate P (the tensor-core kernel computes it)
    int skip_P;
    // CTA kernels, shapes beyond shared memory: global-memory home of the matrices (nullptr: shared memory)
    void *workspace;
    // shared-model kernel: the closed loop it runs (kind 0: none)
    LoopDev loop;
};

template <typename T> struct Pair;
template <> struct Pair<double> { using type = double2; };
template <> struct Pair<float> { using type = float2; };

__device__ __forceinline__ double rsqrt_(double v) { return rsqrt(v); }
__device__ __forceinline__ float rsqrt_(float v) { return rsqrtf(v); }
__device__ __forceinline__ void sincos_(double a, double *s, double *c) { sincos(a, s, c); }
__device__ __forceinline__ void sincos_(float a, float *s, float *c) { sincosf(a, s, c); }
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

// Element i of an array the compiler keeps in registers, for an index that is only known at run
// time but is UNIFORM over the warp (a position in the working set): a switch over static
// indices (a jump table, one taken branch) instead of spilling the array to local memory.
template <typename T, int NP>
__device__ __forceinline__ T reg_get(const T (&a)[NP], int i) {
    static_assert(NP <= 64, "reg_get covers 64 entries");
    switch (i) {
        case 0: if (0 < NP) return a[0 < NP ? 0 : 0]; break;
        case 1: if (1 < NP) return a[1 < NP ? 1 : 0]; break;
        case 2: if (2 < NP) return a[2 < NP ? 2 : 0]; break;
        case 3: if (3 < NP) return a[3 < NP ? 3 : 0]; break;
        case 4: if (4 < NP) return a[4 < NP ? 4 : 0]; break;
        case 5: if (5 < NP) return a[5 < NP ? 5 : 0]; break;
        case 6: if (6 < NP) return a[6 < NP ? 6 : 0]; break;
        case 7: if (7 < NP) return a[7 < NP ? 7 : 0]; break;
        case 8: if (8 < NP) return a[8 < NP ? 8 : 0]; break;
        case 9: if (9 < NP) return a[9 < NP ? 9 : 0]; break;
        case 10: if (10 < NP) return a[10 < NP ? 10 : 0]; break;
        case 11: if (11 < NP) return a[11 < NP ? 11 : 0]; break;
        case 12: if (12 < NP) return a[12 < NP ? 12 : 0]; break;
        case 13: if (13 < NP) return a[13 < NP ? 13 : 0]; break;
        case 14: if (14 < NP) return a[14 < NP ? 14 : 0]; break;
        case 15: if (15 < NP) return a[15 < NP ? 15 : 0]; break;
        case 16: if (16 < NP) return a[16 < NP ? 16 : 0]; break;
        case 17: if (17 < NP) return a[17 < NP ? 17 : 0]; break;
        case 18: if (18 < NP) return a[18 < NP ? 18 : 0]; break;
        case 19: if (19 < NP) return a[19 < NP ? 19 : 0]; break;
        case 20: if (20 < NP) return a[20 < NP ? 20 : 0]; break;
        case 21: if (21 < NP) return a[21 < NP ? 21 : 0]; break;
        case 22: if (22 < NP) return a[22 < NP ? 22 : 0]; break;
        case 23: if (23 < NP) return a[23 < NP ? 23 : 0]; break;
        case 24: if (24 < NP) return a[24 < NP ? 24 : 0]; break;
        case 25: if (25 < NP) return a[25 < NP ? 25 : 0]; break;
        case 26: if (26 < NP) return a[26 < NP ? 26 : 0]; break;
        case 27: if (27 < NP) return a[27 < NP ? 27 : 0]; break;
        case 28: if (28 < NP) return a[28 < NP ? 28 : 0]; break;
        case 29: if (29 < NP) return a[29 < NP ? 29 : 0]; break;
        case 30: if (30 < NP) return a[30 < NP ? 30 : 0]; break;
        case 31: if (31 < NP) return a[31 < NP ? 31 : 0]; break;
        case 32: if (32 < NP) return a[32 < NP ? 32 : 0]; break;
        case 33: if (33 < NP) return a[33 < NP ? 33 : 0]; break;
        case 34: if (34 < NP) return a[34 < NP ? 34 : 0]; break;
        case 35: if (35 < NP) return a[35 < NP ? 35 : 0]; break;
        case 36: if (36 < NP) return a[36 < NP ? 36 : 0]; break;
        case 37: if (37 < NP) return a[37 < NP ? 37 : 0]; break;
        case 38: if (38 < NP) return a[38 < NP ? 38 : 0]; break;
        case 39: if (39 < NP) return a[39 < NP ? 39 : 0]; break;
        case 40: if (40 < NP) return a[40 < NP ? 40 : 0]; break;
        case 41: if (41 < NP) return a[41 < NP ? 41 : 0]; break;
        case 42: if (42 < NP) return a[42 < NP ? 42 : 0]; break;
        case 43: if (43 < NP) return a[43 < NP ? 43 : 0]; break;
        case 44: if (44 < NP) return a[44 < NP ? 44 : 0]; break;
        case 45: if (45 < NP) return a[45 < NP ? 45 : 0]; break;
        case 46: if (46 < NP) return a[46 < NP ? 46 : 0]; break;
        case 47: if (47 < NP) return a[47 < NP ? 47 : 0]; break;
        case 48: if (48 < NP) return a[48 < NP ? 48 : 0]; break;
        case 49: if (49 < NP) return a[49 < NP ? 49 : 0]; break;
        case 50: if (50 < NP) return a[50 < NP ? 50 : 0]; break;
        case 51: if (51 < NP) return a[51 < NP ? 51 : 0]; break;
        case 52: if (52 < NP) return a[52 < NP ? 52 : 0]; break;
        case 53: if (53 < NP) return a[53 < NP ? 53 : 0]; break;
        case 54: if (54 < NP) return a[54 < NP ? 54 : 0]; break;
        case 55: if (55 < NP) return a[55 < NP ? 55 : 0]; break;
        case 56: if (56 < NP) return a[56 < NP ? 56 : 0]; break;
        case 57: if (57 < NP) return a[57 < NP ? 57 : 0]; break;
        case 58: if (58 < NP) return a[58 < NP ? 58 : 0]; break;
        case 59: if (59 < NP) return a[59 < NP ? 59 : 0]; break;
        case 60: if (60 < NP) return a[60 < NP ? 60 : 0]; break;
        case 61: if (61 < NP) return a[61 < NP ? 61 : 0]; break;
        case 62: if (62 < NP) return a[62 < NP ? 62 : 0]; break;
        case 63: if (63 < NP) return a[63 < NP ? 63 : 0]; break;
    }
    return a[0];
}
template <typename T, int NP>
__device__ __forceinline__ void reg_sub(T (&a)[NP], int i, T v) {
    switch (i) {
        case 0: if (0 < NP) a[0 < NP ? 0 : 0] -= v; break;
        case 1: if (1 < NP) a[1 < NP ? 1 : 0] -= v; break;
        case 2: if (2 < NP) a[2 < NP ? 2 : 0] -= v; break;
        case 3: if (3 < NP) a[3 < NP ? 3 : 0] -= v; break;
        case 4: if (4 < NP) a[4 < NP ? 4 : 0] -= v; break;
        case 5: if (5 < NP) a[5 < NP ? 5 : 0] -= v; break;
        case 6: if (6 < NP) a[6 < NP ? 6 : 0] -= v; break;
        case 7: if (7 < NP) a[7 < NP ? 7 : 0] -= v; break;
        case 8: if (8 < NP) a[8 < NP ? 8 : 0] -= v; break;
        case 9: if (9 < NP) a[9 < NP ? 9 : 0] -= v; break;
        case 10: if (10 < NP) a[10 < NP ? 10 : 0] -= v; break;
        case 11: if (11 < NP) a[11 < NP ? 11 : 0] -= v; break;
        case 12: if (12 < NP) a[12 < NP ? 12 : 0] -= v; break;
        case 13: if (13 < NP) a[13 < NP ? 13 : 0] -= v; break;
        case 14: if (14 < NP) a[14 < NP ? 14 : 0] -= v; break;
        case 15: if (15 < NP) a[15 < NP ? 15 : 0] -= v; break;
        case 16: if (16 < NP) a[16 < NP ? 16 : 0] -= v; break;
        case 17: if (17 < NP) a[17 < NP ? 17 : 0] -= v; break;
        case 18: if (18 < NP) a[18 < NP ? 18 : 0] -= v; break;
        case 19: if (19 < NP) a[19 < NP ? 19 : 0] -= v; break;
        case 20: if (20 < NP) a[20 < NP ? 20 : 0] -= v; break;
        case 21: if (21 < NP) a[21 < NP ? 21 : 0] -= v; break;
        case 22: if (22 < NP) a[22 < NP ? 22 : 0] -= v; break;
        case 23: if (23 < NP) a[23 < NP ? 23 : 0] -= v; break;
        case 24: if (24 < NP) a[24 < NP ? 24 : 0] -= v; break;
        case 25: if (25 < NP) a[25 < NP ? 25 : 0] -= v; break;
        case 26: if (26 < NP) a[26 < NP ? 26 : 0] -= v; break;
        case 27: if (27 < NP) a[27 < NP ? 27 : 0] -= v; break;
        case 28: if (28 < NP) a[28 < NP ? 28 : 0] -= v; break;
        case 29: if (29 < NP) a[29 < NP ? 29 : 0] -= v; break;
        case 30: if (30 < NP) a[30 < NP ? 30 : 0] -= v; break;
        case 31: if (31 < NP) a[31 < NP ? 31 : 0] -= v; break;
        case 32: if (32 < NP) a[32 < NP ? 32 : 0] -= v; break;
        case 33: if (33 < NP) a[33 < NP ? 33 : 0] -= v; break;
        case 34: if (34 < NP) a[34 < NP ? 34 : 0] -= v; break;
        case 35: if (35 < NP) a[35 < NP ? 35 : 0] -= v; break;
        case 36: if (36 < NP) a[36 < NP ? 36 : 0] -= v; break;
        case 37: if (37 < NP) a[37 < NP ? 37 : 0] -= v; break;
        case 38: if (38 < NP) a[38 < NP ? 38 : 0] -= v; break;
        case 39: if (39 < NP) a[39 < NP ? 39 : 0] -= v; break;
        case 40: if (40 < NP) a[40 < NP ? 40 : 0] -= v; break;
        case 41: if (41 < NP) a[41 < NP ? 41 : 0] -= v; break;
        case 42: if (42 < NP) a[42 < NP ? 42 : 0] -= v; break;
        case 43: if (43 < NP) a[43 < NP ? 43 : 0] -= v; break;
        case 44: if (44 < NP) a[44 < NP ? 44 : 0] -= v; break;
        case 45: if (45 < NP) a[45 < NP ? 45 : 0] -= v; break;
        case 46: if (46 < NP) a[46 < NP ? 46 : 0] -= v; break;
        case 47: if (47 < NP) a[47 < NP ? 47 : 0] -= v; break;
        case 48: if (48 < NP) a[48 < NP ? 48 : 0] -= v; break;
        case 49: if (49 < NP) a[49 < NP ? 49 : 0] -= v; break;
        case 50: if (50 < NP) a[50 < NP ? 50 : 0] -= v; break;
        case 51: if (51 < NP) a[51 < NP ? 51 : 0] -= v; break;
        case 52: if (52 < NP) a[52 < NP ? 52 : 0] -= v; break;
        case 53: if (53 < NP) a[53 < NP ? 53 : 0] -= v; break;
        case 54: if (54 < NP) a[54 < NP ? 54 : 0] -= v; break;
        case 55: if (55 < NP) a[55 < NP ? 55 : 0] -= v; break;
        case 56: if (56 < NP) a[56 < NP ? 56 : 0] -= v; break;
        case 57: if (57 < NP) a[57 < NP ? 57 : 0] -= v; break;
        case 58: if (58 < NP) a[58 < NP ? 58 : 0] -= v; break;
        case 59: if (59 < NP) a[59 < NP ? 59 : 0] -= v; break;
        case 60: if (60 < NP) a[60 < NP ? 60 : 0] -= v; break;
        case 61: if (61 < NP) a[61 < NP ? 61 : 0] -= v; break;
        case 62: if (62 < NP) a[62 < NP ? 62 : 0] -= v; break;
        case 63: if (63 < NP) a[63 < NP ? 63 : 0] -= v; break;
    }
}

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
