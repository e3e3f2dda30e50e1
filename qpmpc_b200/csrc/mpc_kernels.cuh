// mpc_kernels.cuh -- fused condense + dual active-set QP kernel (sm_100a).
//
// One group of NP lanes (NP = 8, 16 or 32, the padded number of decision
// variables n = N*nu) owns one MPC instance; a warp carries 32/NP instances.
// Lane l is variable l and owns row l of every n x n matrix of the instance
// in REGISTERS; it also owns MR constraint rows (l, l+NP, ...).  Lanes talk
// through a small shared-memory region per instance and warp shuffles; there
// is no block-level synchronisation after the inputs have been staged.
//
// Phases (reference citations relative to /root/reference):
//   0  stage A,B,C,D,e,x0,goal,targets of the CTA's instances into shared
//      memory: one 1-D bulk TMA copy per operand (cp.async.bulk + mbarrier).
//   A  condensing, qpmpc/mpc_qp.py:53-105 and :139-149.  Lane l rolls column
//      l of psi_k and the free response phi_k x0 in registers (nx = 2, 3, 4
//      compiled; other nx go through shared memory), emits column l of every
//      G row and h, accumulates row l of P and q_l.
//   B  Cholesky P = L L' (row per lane, columns published in shared memory),
//      then forward substitutions with L, all lane-local: row l of J = L^-T,
//      t = J'q, x = -J t, the owned rows of M = G J; violations G x - h.
//   C  Goldfarb-Idnani dual active-set iteration on (J, M, R, R^-1) -- the
//      algorithm of the quadprog backend behind qpsolvers.solve_problem
//      (qpmpc/solve_mpc.py:43) -- with a Householder reflection instead of a
//      Givens sweep when a constraint enters, and R^-1 kept explicitly (its
//      new column is -r/beta, free).  Exact on exit: x solves the KKT system
//      of its active set to rounding error.
//   D  write U (coalesced), status, iterations and optionally multipliers.
#pragma once

#include "mpc_common.cuh"

namespace qpmpc {

template <typename T, int NP, int MR>
struct Lay {
    static constexpr int MP = MR * NP;     // padded constraint rows
    static constexpr int LDG = MP + 1;     // G / M by columns: Gc[c*LDG + row]
    static constexpr int LDL = NP + 2;     // L by columns:     Lc[c*LDL + row]
    static constexpr int LDR = NP + 1;     // R by columns:     Rc[k*LDR + row]
    static constexpr int oG = 0;
    static constexpr int szG = ((NP * LDG + 3) / 4) * 4;
    static constexpr int oH = oG + szG;    // hs[MP]
    static constexpr int oRL = oH + MP;    // psi exchange (A), Lc (B), Rc (C)
    static constexpr int szRL = ((NP * LDL + 3) / 4) * 4;
    static constexpr int oV = oRL + szRL;  // qs, xs, dv, dd, d2 [NP each], sc[8]
    static constexpr int szV = 5 * NP + 8;
    static constexpr int fixed = oV + szV;  // runtime-sized tail follows
    // R^-1 (NP x NP by columns) lives in the G region once M is in registers.
    static_assert(NP * NP <= szG, "R^-1 must fit in the G region");
    static_assert(8 * NP <= szRL, "psi exchange buffers must fit in the L region");
};

// Scratch of the generic (any nx) condensing: psi[2][nx][NP], xbar[2][nx],
// phi[2][nx][nx].
__host__ __device__ inline int generic_condense_elems(int NP, int nx) {
    return (2 * nx * NP + 2 * nx + 2 * nx * nx + 3) / 4 * 4;
}
__host__ __device__ inline bool nx_in_registers(int nx) { return nx >= 2 && nx <= 4; }

// Runtime-sized tail of the per-instance region: generic condensing scratch
// and/or R^-1 when M occupies the G region (MREG = false).
__host__ __device__ inline int tail_elems(int NP, int nx, bool mreg) {
    int a = nx_in_registers(nx) ? 0 : generic_condense_elems(NP, nx);
    int b = mreg ? 0 : NP * NP;
    return a > b ? a : b;
}

// ---------------------------------------------------------------------------
// Phase 0: stage the operands of instances [inst0, inst0 + cnt) into `inbase`.
// ---------------------------------------------------------------------------
template <typename T>  // @phase 0 stage inputs
__device__ __forceinline__ void stage_inputs(const SolveParams &p, T *inbase, int inst0, int cnt,
                                             uint64_t *bar) {
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        fence_barrier_init();
    }
    __syncthreads();
    unsigned tx = 0;
    unsigned tma_mask = 0;
#pragma unroll
    for (int o = 0; o < OP_COUNT; ++o) {
        const OperandView &v = p.op[o];
        if (v.ptr == nullptr) continue;
        const T *src = static_cast<const T *>(v.ptr) + (v.per_instance ? (size_t)inst0 * v.sz : 0);
        unsigned bytes = (unsigned)((v.per_instance ? cnt : 1) * v.sz * (int)sizeof(T));
        if (((reinterpret_cast<uintptr_t>(src) & 15) == 0) && ((bytes & 15) == 0) && bytes > 0) {
            tx += bytes;
            tma_mask |= 1u << o;
        }
    }
    if (threadIdx.x == 0) {
        mbar_arrive_expect_tx(bar, tx);
#pragma unroll
        for (int o = 0; o < OP_COUNT; ++o) {
            if (!(tma_mask >> o & 1)) continue;
            const OperandView &v = p.op[o];
            const T *src = static_cast<const T *>(v.ptr) + (v.per_instance ? (size_t)inst0 * v.sz : 0);
            unsigned bytes = (unsigned)((v.per_instance ? cnt : 1) * v.sz * (int)sizeof(T));
            bulk_g2s(inbase + v.smem_off, src, bytes, bar);
        }
    }
    // Operands a bulk copy cannot take (odd tail counts, unaligned views).
#pragma unroll
    for (int o = 0; o < OP_COUNT; ++o) {
        const OperandView &v = p.op[o];
        if (v.ptr == nullptr || (tma_mask >> o & 1)) continue;
        const T *src = static_cast<const T *>(v.ptr) + (v.per_instance ? (size_t)inst0 * v.sz : 0);
        int count = (v.per_instance ? cnt : 1) * v.sz;
        for (int i = threadIdx.x; i < count; i += blockDim.x) inbase[v.smem_off + i] = src[i];
    }
    __syncthreads();
    mbar_wait(bar, 0);
}

// P row += a * (row vector v[0..NP) in shared memory), v read as pairs.
template <typename T, int NP>
__device__ __forceinline__ void prow_axpy(T (&Prow)[NP], T a, const T *v) {
    using T2 = typename Pair<T>::type;
#pragma unroll
    for (int i = 0; i < NP; i += 2) {
        const T2 w = *reinterpret_cast<const T2 *>(v + i);
        Prow[i] += a * w.x;
        Prow[i + 1] += a * w.y;
    }
}

// ---------------------------------------------------------------------------
// Phase A, register path (nx = NX known at compile time).  Column l of psi_k
// and the free response xb = phi_k x0 never leave registers; psi_k goes
// through the exchange buffer `xch` (2 x NX x NP) only when a cost term needs
// the other lanes' columns.  On exit Gc/hs hold G, h; Prow is row l of P; qj
// is q_l.  With DUMP the MPCQP fields Phi, Psi, phi_last, psi_last are stored.
// ---------------------------------------------------------------------------
template <typename T, int NP, int MR, int NX, bool DUMP>  // @phase A condense (registers)
__device__ __forceinline__ void condense_reg(const SolveParams &p, const T *const (&in)[OP_COUNT], T *Gc,
                                             T *hs, T *xch, int l, T (&Prow)[NP], T &qj, long long inst,
                                             bool valid) {
    using L = Lay<T, NP, MR>;
    const int nu = p.nu, nc = p.nc, N = p.N, n = p.n;
    const T w_t = (T)p.w_t, w_x = (T)p.w_x;
    T psi[NX], xb[NX], Ar[NX * NX];
    T phi[DUMP ? NX * NX : 1];
#pragma unroll
    for (int i = 0; i < NP; ++i) Prow[i] = T(0);
    qj = T(0);
#pragma unroll
    for (int t = 0; t < NX; ++t) {
        psi[t] = T(0);
        xb[t] = in[OP_X0][t];
    }
    if (DUMP) {
#pragma unroll
        for (int t = 0; t < NX * NX; ++t) phi[DUMP ? t : 0] = (t / NX == t % NX) ? T(1) : T(0);
    }
    const int stepA = p.op[OP_A].step;
#pragma unroll
    for (int t = 0; t < NX * NX; ++t) Ar[t] = in[OP_A][t];
    for (int k = 0; k < N; ++k) {
        if (stepA != 0 && k > 0) {
#pragma unroll
            for (int t = 0; t < NX * NX; ++t) Ar[t] = in[OP_A][k * stepA + t];
        }
        const T *Bk = in[OP_B] + k * p.op[OP_B].step;
        const T *Ck = in[OP_C] ? in[OP_C] + k * p.op[OP_C].step : nullptr;
        const T *Dk = in[OP_D] ? in[OP_D] + k * p.op[OP_D].step : nullptr;
        const T *ek = in[OP_E] ? in[OP_E] + k * p.op[OP_E].step : nullptr;
        const int jj = l - k * nu;  // position of this lane's variable inside block k
        const bool inblk = (unsigned)jj < (unsigned)nu;
        // G_k = C_k psi_k + [0 .. D_k .. 0], h_k = e_k - C_k (phi_k x0)   (mpc_qp.py:67-78)
        for (int r = 0; r < nc; ++r) {
            T g = T(0), hv = ek[r];
            if (Ck) {
#pragma unroll
                for (int t = 0; t < NX; ++t) {
                    const T c = Ck[r * NX + t];
                    g += c * psi[t];
                    hv -= c * xb[t];
                }
            }
            if (Dk && inblk) g += Dk[r * nu + jj];
            Gc[l * L::LDG + k * nc + r] = g;
            if (l == 0) hs[k * nc + r] = hv;
        }
        if (DUMP && valid) {
            if (p.Psi && l < n) {
#pragma unroll
                for (int t = 0; t < NX; ++t)
                    static_cast<T *>(p.Psi)[((size_t)inst * N * NX + (size_t)k * NX + t) * n + l] = psi[t];
            }
            if (p.Phi && l == 0) {
#pragma unroll
                for (int t = 0; t < NX * NX; ++t)
                    static_cast<T *>(p.Phi)[((size_t)inst * N * NX + (size_t)k * NX) * NX + t] = phi[DUMP ? t : 0];
            }
        }
        // stage cost: P += w_x psi_k' psi_k, q += w_x psi_k'(phi_k x0 - target_k)
        if (p.has_wx) {
            T *buf = xch + (k & 1) * NX * NP;
#pragma unroll
            for (int t = 0; t < NX; ++t) buf[t * NP + l] = psi[t];
            __syncwarp();
#pragma unroll
            for (int t = 0; t < NX; ++t) prow_axpy<T, NP>(Prow, w_x * psi[t], buf + t * NP);
        }
        if (p.q_wx) {
            const T *tg = in[OP_TGT] + k * NX;
#pragma unroll
            for (int t = 0; t < NX; ++t) qj += (w_x * psi[t]) * (xb[t] - tg[t]);
        }
        // psi_{k+1} = A_k psi_k, block column k := B_k ; xb_{k+1} = A_k xb_k   (mpc_qp.py:88-90)
        T pn[NX], xn[NX];
#pragma unroll
        for (int t = 0; t < NX; ++t) {
            T a = T(0), b = T(0);
#pragma unroll
            for (int s = 0; s < NX; ++s) {
                a += Ar[t * NX + s] * psi[s];
                b += Ar[t * NX + s] * xb[s];
            }
            if (inblk) a = Bk[t * nu + jj];
            pn[t] = a;
            xn[t] = b;
        }
        if (DUMP) {
            T pp[DUMP ? NX * NX : 1];
#pragma unroll
            for (int t = 0; t < NX * NX; ++t) {
                T a = T(0);
#pragma unroll
                for (int s = 0; s < NX; ++s) a += Ar[(t / NX) * NX + s] * phi[DUMP ? s * NX + t % NX : 0];
                pp[DUMP ? t : 0] = a;
            }
#pragma unroll
            for (int t = 0; t < NX * NX; ++t) phi[DUMP ? t : 0] = pp[DUMP ? t : 0];
        }
#pragma unroll
        for (int t = 0; t < NX; ++t) {
            psi[t] = pn[t];
            xb[t] = xn[t];
        }
    }
    // terminal cost: P += w_t psi_N' psi_N, q += w_t psi_N'(phi_N x0 - goal)
    if (p.has_wt) {
        T *buf = xch + (N & 1) * NX * NP;
#pragma unroll
        for (int t = 0; t < NX; ++t) buf[t * NP + l] = psi[t];
        __syncwarp();
#pragma unroll
        for (int t = 0; t < NX; ++t) prow_axpy<T, NP>(Prow, w_t * psi[t], buf + t * NP);
    }
    if (p.q_wt) {
        const T *goal = in[OP_GOAL];
#pragma unroll
        for (int t = 0; t < NX; ++t) qj += (w_t * psi[t]) * (xb[t] - goal[t]);
    }
    // + w_u I on the real variables, identity on the padding.
#pragma unroll
    for (int i = 0; i < NP; ++i)
        if (i == l) Prow[i] += (l < n) ? (T)p.w_u : T(1);
    if (DUMP && valid) {
        if (p.psi_last && l < n) {
#pragma unroll
            for (int t = 0; t < NX; ++t) static_cast<T *>(p.psi_last)[((size_t)inst * NX + t) * n + l] = psi[t];
        }
        if (p.phi_last && l == 0) {
#pragma unroll
            for (int t = 0; t < NX * NX; ++t)
                static_cast<T *>(p.phi_last)[(size_t)inst * NX * NX + t] = phi[DUMP ? t : 0];
        }
    }
    __syncwarp();  // G, h complete; exchange buffer free
}

// ---------------------------------------------------------------------------
// Phase A, generic path (any nx): psi ping-pongs through shared memory.
// `scr` = psi[2][nx][NP], xbar[2][nx], phi[2][nx][nx].
// ---------------------------------------------------------------------------
template <typename T, int NP, int MR, bool DUMP>  // @phase A condense (generic nx)
__device__ __forceinline__ void condense_generic(const SolveParams &p, const T *const (&in)[OP_COUNT], T *Gc,
                                              T *hs, T *scr, int l, T (&Prow)[NP], T &qj, long long inst,
                                              bool valid) {
    using L = Lay<T, NP, MR>;
    const int nx = p.nx, nu = p.nu, nc = p.nc, N = p.N, n = p.n;
    T *psi = scr;
    T *xbar = psi + 2 * nx * NP;
    T *phi = xbar + 2 * nx;
    const T w_t = (T)p.w_t, w_x = (T)p.w_x;
#pragma unroll
    for (int i = 0; i < NP; ++i) Prow[i] = T(0);
    qj = T(0);
    for (int t = 0; t < nx; ++t) psi[t * NP + l] = T(0);
    for (int t = l; t < nx; t += NP) xbar[t] = in[OP_X0][t];
    if (DUMP) {
        for (int t = l; t < nx * nx; t += NP) phi[t] = (t / nx == t % nx) ? T(1) : T(0);
    }
    __syncwarp();
    int cur = 0;
    for (int k = 0; k < N; ++k) {
        const T *Ak = in[OP_A] + k * p.op[OP_A].step;
        const T *Bk = in[OP_B] + k * p.op[OP_B].step;
        const T *Ck = in[OP_C] ? in[OP_C] + k * p.op[OP_C].step : nullptr;
        const T *Dk = in[OP_D] ? in[OP_D] + k * p.op[OP_D].step : nullptr;
        const T *ek = in[OP_E] ? in[OP_E] + k * p.op[OP_E].step : nullptr;
        const T *ps = psi + cur * nx * NP;
        T *pn = psi + (cur ^ 1) * nx * NP;
        const T *xb = xbar + cur * nx;
        T *xn = xbar + (cur ^ 1) * nx;
        const int jj = l - k * nu;
        for (int r = 0; r < nc; ++r) {
            T g = T(0);
            if (Ck)
                for (int t = 0; t < nx; ++t) g += Ck[r * nx + t] * ps[t * NP + l];
            if (Dk && jj >= 0 && jj < nu) g += Dk[r * nu + jj];
            Gc[l * L::LDG + k * nc + r] = g;
        }
        for (int r = l; r < nc; r += NP) {
            T hv = ek[r];
            if (Ck)
                for (int t = 0; t < nx; ++t) hv -= Ck[r * nx + t] * xb[t];
            hs[k * nc + r] = hv;
        }
        if (DUMP && valid) {
            if (p.Psi && l < n)
                for (int t = 0; t < nx; ++t)
                    static_cast<T *>(p.Psi)[((size_t)inst * N * nx + (size_t)k * nx + t) * n + l] = ps[t * NP + l];
            if (p.Phi) {
                const T *ph = phi + cur * nx * nx;
                for (int t = l; t < nx * nx; t += NP)
                    static_cast<T *>(p.Phi)[((size_t)inst * N * nx + (size_t)k * nx) * nx + t] = ph[t];
            }
        }
        if (p.has_wx)
            for (int t = 0; t < nx; ++t) prow_axpy<T, NP>(Prow, w_x * ps[t * NP + l], ps + t * NP);
        if (p.q_wx) {
            const T *tg = in[OP_TGT] + k * nx;
            for (int t = 0; t < nx; ++t) qj += (w_x * ps[t * NP + l]) * (xb[t] - tg[t]);
        }
        for (int t = 0; t < nx; ++t) {
            T acc = T(0);
            for (int s = 0; s < nx; ++s) acc += Ak[t * nx + s] * ps[s * NP + l];
            if (jj >= 0 && jj < nu) acc = Bk[t * nu + jj];
            pn[t * NP + l] = acc;
        }
        for (int t = l; t < nx; t += NP) {
            T acc = T(0);
            for (int s = 0; s < nx; ++s) acc += Ak[t * nx + s] * xb[s];
            xn[t] = acc;
        }
        if (DUMP) {
            const T *ph = phi + cur * nx * nx;
            T *pnx = phi + (cur ^ 1) * nx * nx;
            for (int t = l; t < nx * nx; t += NP) {
                const int r = t / nx, c = t % nx;
                T acc = T(0);
                for (int s = 0; s < nx; ++s) acc += Ak[r * nx + s] * ph[s * nx + c];
                pnx[t] = acc;
            }
        }
        __syncwarp();
        cur ^= 1;
    }
    const T *ps = psi + cur * nx * NP;
    const T *xb = xbar + cur * nx;
    if (p.has_wt)
        for (int t = 0; t < nx; ++t) prow_axpy<T, NP>(Prow, w_t * ps[t * NP + l], ps + t * NP);
    if (p.q_wt) {
        const T *goal = in[OP_GOAL];
        for (int t = 0; t < nx; ++t) qj += (w_t * ps[t * NP + l]) * (xb[t] - goal[t]);
    }
#pragma unroll
    for (int i = 0; i < NP; ++i)
        if (i == l) Prow[i] += (l < n) ? (T)p.w_u : T(1);
    if (DUMP && valid) {
        if (p.psi_last && l < n)
            for (int t = 0; t < nx; ++t) static_cast<T *>(p.psi_last)[((size_t)inst * nx + t) * n + l] = ps[t * NP + l];
        if (p.phi_last) {
            const T *ph = phi + cur * nx * nx;
            for (int t = l; t < nx * nx; t += NP) static_cast<T *>(p.phi_last)[(size_t)inst * nx * nx + t] = ph[t];
        }
    }
    __syncwarp();
}

template <typename T, int NP, int MR, bool DUMP>
__device__ __forceinline__ void condense_dispatch(const SolveParams &p, const T *const (&in)[OP_COUNT], T *Gc,
                                                  T *hs, T *xch, T *tail, int l, T (&Prow)[NP], T &qj,
                                                  long long inst, bool valid) {
    switch (p.nx) {
        case 2: condense_reg<T, NP, MR, 2, DUMP>(p, in, Gc, hs, xch, l, Prow, qj, inst, valid); break;
        case 3: condense_reg<T, NP, MR, 3, DUMP>(p, in, Gc, hs, xch, l, Prow, qj, inst, valid); break;
        case 4: condense_reg<T, NP, MR, 4, DUMP>(p, in, Gc, hs, xch, l, Prow, qj, inst, valid); break;
        default: condense_generic<T, NP, MR, DUMP>(p, in, Gc, hs, tail, l, Prow, qj, inst, valid); break;
    }
}

// ---------------------------------------------------------------------------
// Forward substitution L y = rhs for two right-hand sides held in registers,
// L by columns in shared memory (Lc), 1/L_kk in dv.  Lane-local: every lane
// solves its own systems; all lanes read the same L entries (broadcast).
// ---------------------------------------------------------------------------
template <typename T, int NP, int LDL>
__device__ __forceinline__ void fsolve2(const T *Lc, const T *dv, T (&a)[NP], T (&b)[NP]) {
    using T2 = typename Pair<T>::type;
#pragma unroll
    for (int k = 0; k < NP; ++k) {
        const T dk = dv[k];
        const T ak = a[k] * dk, bk = b[k] * dk;
        a[k] = ak;
        b[k] = bk;
        if (((k + 1) & 1) && k + 1 < NP) {
            const T lv = Lc[k * LDL + k + 1];
            a[k + 1] -= lv * ak;
            b[k + 1] -= lv * bk;
        }
#pragma unroll
        for (int c = (k + 2) & ~1; c < NP; c += 2) {
            const T2 lv = *reinterpret_cast<const T2 *>(Lc + k * LDL + c);
            a[c] -= lv.x * ak;
            a[c + 1] -= lv.y * ak;
            b[c] -= lv.x * bk;
            b[c + 1] -= lv.y * bk;
        }
    }
}

// Sortable 64-bit key of a positive score and a row index < 128: larger score
// wins, ties go to the lower index.  0 means "no candidate".
__device__ __forceinline__ unsigned long long score_key(double s, int idx) {
    return ((unsigned long long)__double_as_longlong(s) & ~127ull) | (unsigned)(127 - idx);
}
__device__ __forceinline__ unsigned long long score_key(float s, int idx) {
    return ((unsigned long long)__float_as_uint(s) << 32) | (unsigned)(127 - idx);
}

// ---------------------------------------------------------------------------
// The fused kernel.  MREG: rows of M = G J live in registers (NP <= 16);
// otherwise M overwrites G in shared memory.
// ---------------------------------------------------------------------------
template <typename T, int NP, int MR, bool MREG>  // @phase kernel prologue
__global__ void __launch_bounds__(128, (NP <= 16 && MREG) ? 3 : 1) mpc_solve_kernel(const SolveParams p) {
    using L = Lay<T, NP, MR>;
    using T2 = typename Pair<T>::type;
    constexpr int IPW = 32 / NP;  // instances per warp
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw);
    T *work = reinterpret_cast<T *>(smem_raw + 16);

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wpc = blockDim.x >> 5;
    const int ipc = IPW * wpc;
    const int sub = lane / NP;
    const int l = lane % NP;
    const int seg_shift = sub * NP;
    const int iic = warp * IPW + sub;  // instance slot inside the CTA
    const int inst0 = blockIdx.x * ipc;
    const int cnt = min(ipc, p.batch - inst0);
    const long long inst = (long long)inst0 + iic;
    const bool valid = iic < cnt;
    const int n = p.n, m = p.m;

    T *inbase = work + (size_t)ipc * p.inst_stride;
    stage_inputs<T>(p, inbase, inst0, cnt, bar);

    T *wk = work + (size_t)iic * p.inst_stride;
    T *Gc = wk + L::oG;
    T *hs = wk + L::oH;
    T *Lc = wk + L::oRL;
    T *Rc = wk + L::oRL;
    T *qs = wk + L::oV;
    T *xs = qs + NP;
    T *dv = qs + 2 * NP;
    T *dd = qs + 3 * NP;
    T *d2 = qs + 4 * NP;
    T *sc = qs + 5 * NP;
    T *tail = wk + L::fixed;
    T *Ri = MREG ? Gc : tail;  // R^-1 by columns: Ri[k*NP + row]

    const T *in[OP_COUNT];
#pragma unroll
    for (int o = 0; o < OP_COUNT; ++o) {
        const OperandView &v = p.op[o];
        // Invalid tail slots read instance 0 of the CTA: defined data, results discarded.
        in[o] = v.ptr ? inbase + v.smem_off + (v.per_instance ? (valid ? iic : 0) * v.sz : 0) : nullptr;
    }

    // ---- phase A -----------------------------------------------------------
    T Prow[NP];
    T qj;
    condense_dispatch<T, NP, MR, false>(p, in, Gc, hs, Lc, tail, l, Prow, qj, inst, valid);

    // ---- phase B: Cholesky, row l of L in Prow, columns published in Lc -----  // @phase B cholesky
    bool spd = true;
    qs[l] = qj;
#pragma unroll
    for (int c = 0; c < NP; ++c) {
        const T piv = __shfl_sync(FULL_MASK, Prow[c], c, NP);
        spd = spd && (piv > T(0));
        const T inv = rsqrt_(piv);
        const T lc = Prow[c] * inv;
        Lc[c * L::LDL + l] = lc;
        if (l == c) dv[c] = inv;
        __syncwarp();
        if (((c + 1) & 1) && c + 1 < NP) Prow[c + 1] -= lc * Lc[c * L::LDL + c + 1];
#pragma unroll
        for (int i = (c + 2) & ~1; i < NP; i += 2) {
            const T2 v = *reinterpret_cast<const T2 *>(Lc + c * L::LDL + i);
            Prow[i] -= lc * v.x;
            Prow[i + 1] -= lc * v.y;
        }
    }
    // Row l of J = L^-T is column l of L^-1: solve L y = e_l; t = L^-1 q = J'q
    // with the same sweep; x = -J t.  // @phase B J=L^-T, x=-P^-1 q
    T Jrow[NP];
    T x;
    {
        T tq[NP];
#pragma unroll
        for (int c = 0; c < NP; c += 2) {
            const T2 v = *reinterpret_cast<const T2 *>(qs + c);
            tq[c] = v.x;
            tq[c + 1] = v.y;
            Jrow[c] = (c == l) ? T(1) : T(0);
            Jrow[c + 1] = (c + 1 == l) ? T(1) : T(0);
        }
        fsolve2<T, NP, L::LDL>(Lc, dv, Jrow, tq);
        T x0 = T(0), x1 = T(0);
#pragma unroll
        for (int c = 0; c < NP; c += 2) {
            x0 -= Jrow[c] * tq[c];
            x1 -= Jrow[c + 1] * tq[c + 1];
        }
        x = x0 + x1;
    }
    xs[l] = x;
    __syncwarp();

    // Owned rows of M = G J: solve L m' = g' per row; violations, row norms.  // @phase B M=GJ, violations
    T Mrow[MREG ? MR : 1][NP];
    T viol[MR], mn2[MR], vtol[MR], ginv[MR];
    bool rowvalid[MR];
#pragma unroll
    for (int s0 = 0; s0 < MR; s0 += 2) {
        T acc[2][NP];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int s = s0 + u;
            const int row = l + s * NP;
            rowvalid[s] = row < m;
            T vi0 = T(0), vi1 = T(0), g2 = T(0);
#pragma unroll
            for (int k = 0; k < NP; k += 2) {
                const T ga = rowvalid[s] ? Gc[k * L::LDG + row] : T(0);
                const T gb = rowvalid[s] ? Gc[(k + 1) * L::LDG + row] : T(0);
                const T2 xv = *reinterpret_cast<const T2 *>(xs + k);
                acc[u][k] = ga;
                acc[u][k + 1] = gb;
                g2 += ga * ga;
                g2 += gb * gb;
                vi0 += ga * xv.x;
                vi1 += gb * xv.y;
            }
            // Tolerance of the violation test: eps * (max(1, |h_i|) + |G_i|).
            const T hi = rowvalid[s] ? hs[row] : T(0);
            viol[s] = rowvalid[s] ? (vi0 + vi1) - hi : T(-1);
            vtol[s] = Num<T>::viol_eps * (fmax(T(1), abs_(hi)) + sqrt_(g2));
            ginv[s] = g2 > T(0) ? rsqrt_(g2) : T(1e30);
        }
        fsolve2<T, NP, L::LDL>(Lc, dv, acc[0], acc[1]);
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int s = s0 + u;
            T m0 = T(0), m1 = T(0);
#pragma unroll
            for (int c = 0; c < NP; c += 2) {
                m0 += acc[u][c] * acc[u][c];
                m1 += acc[u][c + 1] * acc[u][c + 1];
            }
            mn2[s] = m0 + m1;
            if (MREG) {
#pragma unroll
                for (int c = 0; c < NP; ++c) Mrow[MREG ? s : 0][c] = acc[u][c];
            } else {
                // In place: this lane is the only reader and writer of its rows of G.
                const int row = l + s * NP;
#pragma unroll
                for (int c = 0; c < NP; ++c) Gc[c * L::LDG + row] = acc[u][c];
            }
        }
    }
    auto mget = [&](int s, int c) -> T { return MREG ? Mrow[MREG ? s : 0][c] : Gc[c * L::LDG + l + s * NP]; };
    auto mset = [&](int s, int c, T v) {
        if (MREG)
            Mrow[MREG ? s : 0][c] = v;
        else
            Gc[c * L::LDG + l + s * NP] = v;
    };
    __syncwarp();  // G (when MREG) and Lc are dead from here on: R^-1 and R take their place

    // ---- phase C: dual active-set iteration ---------------------------------  // @phase C select row (step 1)
    const int max_iter = p.max_iter;
    int na = 0, it = 0;
    int st = spd ? 0 : 3;
    bool done = !valid || st != 0 || m == 0;
    bool cont = false;
    int pidx = 0;
    T lam = T(0), lamp = T(0);
    int aidx = -1;
    unsigned actbits = 0;
    const T INF = Num<T>::inf();

    while (true) {
        {
            // step 1: most violated inactive row, relative to its norm
            const bool sel = !done && !cont;
            unsigned long long key = 0ull;
#pragma unroll
            for (int s = 0; s < MR; ++s) {
                if (sel && rowvalid[s] && !((actbits >> s) & 1) && viol[s] > vtol[s]) {
                    const unsigned long long ks = score_key(viol[s] * ginv[s], l + s * NP);
                    key = ks > key ? ks : key;
                }
            }
#pragma unroll
            for (int off = NP / 2; off > 0; off >>= 1) {
                const unsigned long long o = __shfl_xor_sync(FULL_MASK, key, off, NP);
                key = o > key ? o : key;
            }
            if (sel) {
                if (key == 0ull) {
                    done = true;  // primal feasible: optimal
                } else {
                    pidx = 127 - (int)(key & 127ull);
                    lamp = T(0);
                }
            }
        }
        if (__all_sync(FULL_MASK, done)) break;
        bool act = !done;
        if (act) {
            ++it;
            if (it > max_iter) {
                st = 1;
                done = true;
                act = false;
            }
        }
        // d = J' n_p = -(row p of M): dd holds d, d2 its part beyond the active columns.  // @phase C publish d
        const int owner = pidx % NP, pslot = pidx / NP;
        T dl;
        if (MREG) {
            if (act && l == owner) {
#pragma unroll
                for (int s = 0; s < MR; ++s) {
                    if (s == pslot) {
#pragma unroll
                        for (int c = 0; c < NP; c += 2) {
                            T2 v;
                            v.x = -Mrow[MREG ? s : 0][c];
                            v.y = -Mrow[MREG ? s : 0][c + 1];
                            *reinterpret_cast<T2 *>(dd + c) = v;
                        }
                        sc[0] = viol[s];
                        sc[1] = mn2[s];
                    }
                }
            }
            __syncwarp();
            dl = dd[l];
        } else {
            dl = -Gc[l * L::LDG + pidx];
            dd[l] = dl;
            if (act && l == owner) {
#pragma unroll
                for (int s = 0; s < MR; ++s) {
                    if (s == pslot) {
                        sc[0] = viol[s];
                        sc[1] = mn2[s];
                    }
                }
            }
        }
        d2[l] = (l >= na) ? dl : T(0);
        __syncwarp();
        // z = J2 d2 (this lane's component), G z (owned rows), |d2|^2  // @phase C z, Gz
        T z = T(0), z1 = T(0), a2 = T(0), a21 = T(0);
        T gz[MR];
#pragma unroll
        for (int s = 0; s < MR; ++s) gz[s] = T(0);
#pragma unroll
        for (int c = 0; c < NP; c += 2) {
            const T2 v = *reinterpret_cast<const T2 *>(d2 + c);
            z += Jrow[c] * v.x;
            z1 += Jrow[c + 1] * v.y;
            a2 += v.x * v.x;
            a21 += v.y * v.y;
#pragma unroll
            for (int s = 0; s < MR; ++s) {
                gz[s] += mget(s, c) * v.x;
                gz[s] += mget(s, c + 1) * v.y;
            }
        }
        z += z1;
        a2 += a21;
        // r = R^-1 d1 (component l on lane l < na)  // @phase C r=R^-1 d
        T rv = T(0);
        {
            const int namax = __reduce_max_sync(FULL_MASK, act ? na : 0);
            for (int k = 0; k < namax; ++k) {
                const T dk = dd[k];
                const T ri = Ri[k * NP + l];
                if (l <= k && k < na) rv += ri * dk;
            }
        }
        // step lengths  // @phase C step length, move
        const T cand = (act && l < na && rv > T(0)) ? lam * rcp_(rv) : INF;
        T t1 = cand;
#pragma unroll
        for (int off = NP / 2; off > 0; off >>= 1) t1 = fmin(t1, __shfl_xor_sync(FULL_MASK, t1, off, NP));
        const unsigned bal = __ballot_sync(FULL_MASK, cand == t1 && cand < INF);
        const unsigned segbits = (NP == 32) ? bal : ((bal >> seg_shift) & ((1u << (NP & 31)) - 1u));
        const int lidx = segbits ? (__ffs(segbits) - 1) : 0;
        const T violp = sc[0], dn2 = sc[1];
        const bool zzero = !(a2 > Num<T>::dep_eps * dn2);
        const T t2 = zzero ? INF : violp * rcp_(a2);
        if (act && t1 == INF && t2 == INF) {
            st = 2;  // infeasible
            done = true;
            act = false;
        }
        const T t = fmin(t1, t2);
        if (act) {
            if (!zzero) {
                x += t * z;
#pragma unroll
                for (int s = 0; s < MR; ++s) viol[s] += t * gz[s];
            }
            if (l < na) lam -= t * rv;
            lamp += t;
        }
        const bool full = act && !zzero && t2 <= t1;
        const bool part = act && !full;

        if (__any_sync(FULL_MASK, full)) {  // @phase C add constraint (Householder)
            // Constraint p enters: reflect d2 onto its first entry.  H = I - tau v v',
            // v = d2 - beta e_na, applied to columns >= na of J and M.
            const T dna = d2[na < NP ? na : NP - 1];
            const T ainv = rsqrt_(a2);
            const T alpha = a2 * ainv;
            const T beta = (dna > T(0)) ? -alpha : alpha;
            const T binv = (dna > T(0)) ? -ainv : ainv;
            const T tau = full ? rcp_(a2 - beta * dna) : T(0);
            __syncwarp();
            if (full && l == na) d2[l] = dna - beta;  // d2 becomes v
            __syncwarp();
            T dj = T(0), dj1 = T(0);
            T dm[MR];
#pragma unroll
            for (int s = 0; s < MR; ++s) dm[s] = T(0);
#pragma unroll
            for (int c = 0; c < NP; c += 2) {
                const T2 v = *reinterpret_cast<const T2 *>(d2 + c);
                dj += Jrow[c] * v.x;
                dj1 += Jrow[c + 1] * v.y;
#pragma unroll
                for (int s = 0; s < MR; ++s) {
                    dm[s] += mget(s, c) * v.x;
                    dm[s] += mget(s, c + 1) * v.y;
                }
            }
            dj = (dj + dj1) * tau;
#pragma unroll
            for (int s = 0; s < MR; ++s) dm[s] *= tau;
#pragma unroll
            for (int c = 0; c < NP; c += 2) {
                const T2 v = *reinterpret_cast<const T2 *>(d2 + c);
                Jrow[c] -= dj * v.x;
                Jrow[c + 1] -= dj * v.y;
#pragma unroll
                for (int s = 0; s < MR; ++s) {
                    mset(s, c, mget(s, c) - dm[s] * v.x);
                    mset(s, c + 1, mget(s, c + 1) - dm[s] * v.y);
                }
            }
            if (full) {
                // new column of R: [d1; beta]; of R^-1: [-r / beta; 1 / beta]
                if (l < na) {
                    Rc[na * L::LDR + l] = dl;
                    Ri[na * NP + l] = -rv * binv;
                }
                if (l == na) {
                    Rc[na * L::LDR + na] = beta;
                    Ri[na * NP + na] = binv;
                    lam = lamp;
                    aidx = pidx;
                }
                if (l == owner) actbits |= 1u << pslot;
                ++na;
                cont = false;
            }
        }
        if (__any_sync(FULL_MASK, part)) {  // @phase C drop constraint (Givens)
            // Constraint at active position lidx leaves; p stays the candidate.
            const int cidx = __shfl_sync(FULL_MASK, aidx, lidx, NP);
            if (part && l == cidx % NP) actbits &= ~(1u << (cidx / NP));
            const T lam_n = __shfl_down_sync(FULL_MASK, lam, 1, NP);
            const int aidx_n = __shfl_down_sync(FULL_MASK, aidx, 1, NP);
            const int nan_ = na - 1;
            const bool mover = part && l >= lidx && l < nan_;
            if (mover) {
                lam = lam_n;
                aidx = aidx_n;
            }
            // shift columns lidx+1.. of R one to the left
            const int namax = __reduce_max_sync(FULL_MASK, part ? na : 0);
            for (int row = 0; row < namax; ++row) {
                T v = T(0);
                const bool mv = mover && row <= l + 1;
                if (mv) v = Rc[(l + 1) * L::LDR + row];
                __syncwarp();
                if (mv) Rc[l * L::LDR + row] = v;
            }
            __syncwarp();
            // Givens rotations of rows (j, j+1) of R restore the triangle; the
            // same rotations act on columns (j, j+1) of J and M.
#pragma unroll
            for (int j = 0; j < NP - 1; ++j) {
                const bool rot = part && j >= lidx && j < nan_;
                if (!__any_sync(FULL_MASK, rot)) continue;
                T a = T(1), b = T(0);
                if (rot) {
                    a = Rc[j * L::LDR + j];
                    b = Rc[j * L::LDR + j + 1];
                }
                const T h2 = a * a + b * b;
                const T hinv = h2 > T(0) ? rsqrt_(h2) : T(0);
                const T cs = h2 > T(0) ? a * hinv : T(1);
                const T sn = b * hinv;
                if (rot && l >= j && l < nan_) {
                    const T u = Rc[l * L::LDR + j], v = Rc[l * L::LDR + j + 1];
                    Rc[l * L::LDR + j] = cs * u + sn * v;
                    Rc[l * L::LDR + j + 1] = cs * v - sn * u;
                }
                if (rot) {
                    const T u = Jrow[j], v = Jrow[j + 1];
                    Jrow[j] = cs * u + sn * v;
                    Jrow[j + 1] = cs * v - sn * u;
#pragma unroll
                    for (int s = 0; s < MR; ++s) {
                        const T mu = mget(s, j), mv = mget(s, j + 1);
                        mset(s, j, cs * mu + sn * mv);
                        mset(s, j + 1, cs * mv - sn * mu);
                    }
                }
                __syncwarp();
            }
            // R^-1 of the reduced factor: lane j solves R y = e_j into column j.
            if (part && l < nan_) {
                const int j = l;
                Ri[j * NP + j] = T(1) / Rc[j * L::LDR + j];
                for (int i = j - 1; i >= 0; --i) {
                    T s = T(0);
                    for (int k = i + 1; k <= j; ++k) s += Rc[k * L::LDR + i] * Ri[j * NP + k];
                    Ri[j * NP + i] = -s / Rc[i * L::LDR + i];
                }
            }
            if (part) {
                na = nan_;
                cont = true;
            }
        }
        __syncwarp();
    }

    // ---- phase D: outputs ---------------------------------------------------  // @phase D outputs
    if (valid) {
        if (l < n) static_cast<T *>(p.U)[(size_t)inst * n + l] = (st == 0) ? x : Num<T>::nan();
        if (l == 0) {
            p.status[inst] = st;
            if (p.iters) p.iters[inst] = it;
        }
        if (p.Z) {
            T *Zb = static_cast<T *>(p.Z) + (size_t)inst * m;
            for (int r = l; r < m; r += NP) Zb[r] = T(0);
            __syncwarp();
            if (st == 0 && l < na) Zb[aidx] = lam;
        }
    } else if (p.Z) {
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------
// Condense-only kernel: materialises the MPCQP fields for parity checks and
// for the MPCQP host class (qpmpc/mpc_qp.py:28-37).  Same phase-A code as the
// fused kernel.
// ---------------------------------------------------------------------------
template <typename T, int NP, int MR>  // @phase condense-only kernel
__global__ void __launch_bounds__(128) mpc_condense_kernel(const SolveParams p) {
    using L = Lay<T, NP, MR>;
    constexpr int IPW = 32 / NP;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw);
    T *work = reinterpret_cast<T *>(smem_raw + 16);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wpc = blockDim.x >> 5;
    const int ipc = IPW * wpc;
    const int sub = lane / NP, l = lane % NP;
    const int iic = warp * IPW + sub;
    const int inst0 = blockIdx.x * ipc;
    const int cnt = min(ipc, p.batch - inst0);
    const long long inst = (long long)inst0 + iic;
    const bool valid = iic < cnt;
    const int n = p.n, m = p.m;

    T *inbase = work + (size_t)ipc * p.inst_stride;
    stage_inputs<T>(p, inbase, inst0, cnt, bar);
    T *wk = work + (size_t)iic * p.inst_stride;
    T *Gc = wk + L::oG;
    T *hs = wk + L::oH;
    T *xch = wk + L::oRL;
    T *tail = wk + L::fixed;
    const T *in[OP_COUNT];
#pragma unroll
    for (int o = 0; o < OP_COUNT; ++o) {
        const OperandView &v = p.op[o];
        in[o] = v.ptr ? inbase + v.smem_off + (v.per_instance ? (valid ? iic : 0) * v.sz : 0) : nullptr;
    }
    T Prow[NP];
    T qj;
    condense_dispatch<T, NP, MR, true>(p, in, Gc, hs, xch, tail, l, Prow, qj, inst, valid);
    if (!valid) return;
    if (p.P && l < n) {
        T *Pb = static_cast<T *>(p.P) + ((size_t)inst * n + l) * n;
#pragma unroll
        for (int i = 0; i < NP; ++i)
            if (i < n) Pb[i] = Prow[i];
    }
    if (p.q && l < n) static_cast<T *>(p.q)[(size_t)inst * n + l] = qj;
    if (p.G && l < n) {
        T *Gb = static_cast<T *>(p.G) + (size_t)inst * m * n;
        for (int r = 0; r < m; ++r) Gb[(size_t)r * n + l] = Gc[l * L::LDG + r];
    }
    if (p.h) {
        T *hb = static_cast<T *>(p.h) + (size_t)inst * m;
        for (int r = l; r < m; r += NP) hb[r] = hs[r];
    }
}

}  // namespace qpmpc
