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
//   A  condensing, qpmpc/mpc_qp.py:53-105 and :139-149: roll psi_k (lane l
//      holds column l), emit G rows and h, accumulate P row l and q_l.
//   B  Cholesky P = L L' (row per lane), J = L^-T (row per lane),
//      x = -P^-1 q, M = G J (MR rows per lane) and violations G x - h.
//   C  Goldfarb-Idnani dual active-set iteration on (J, M, R) -- the
//      algorithm of the quadprog backend behind qpsolvers.solve_problem
//      (qpmpc/solve_mpc.py:43) -- with a Householder reflection instead of a
//      Givens sweep when a constraint enters.  Exact on exit: x solves the
//      KKT system of its active set to rounding error.
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
    static constexpr int oRL = oH + MP;    // Lc, later Rc
    static constexpr int szRL = ((NP * LDL + 3) / 4) * 4;
    static constexpr int oV = oRL + szRL;  // qs, xs, ts, dfull, d2 [NP each], sc[8]
    static constexpr int szV = 5 * NP + 8;
    static constexpr int oJ = oV + szV;    // psi ping-pong (phase A) / J (phase B)
    static constexpr int fixed = oJ;
};

// Size of the runtime-sized tail region: psi[2][nx][NP], xbar[2][nx],
// phi[2][nx][nx] during condensing, J[NP][NP] afterwards.
__host__ __device__ inline int psi_region_elems(int NP, int nx) {
    int a = 2 * nx * NP + 2 * nx + 2 * nx * nx;
    int b = NP * NP;
    int v = a > b ? a : b;
    return (v + 3) / 4 * 4;
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

// ---------------------------------------------------------------------------
// Phase A: condensing for one instance (NP lanes).  On exit Gc/hs hold G, h,
// Prow is row l of P, qj is q_l, and psi (buffer returned) holds psi_N,
// xbar the free response phi_N x0.
// ---------------------------------------------------------------------------
template <typename T, int NP, int MR, bool DUMP>  // @phase A condense
__device__ __forceinline__ int condense_instance(const SolveParams &p, const T *const (&in)[OP_COUNT],
                                                 T *Gc, T *hs, T *psi, int l, T (&Prow)[NP], T &qj,
                                                 long long inst, bool valid) {
    using L = Lay<T, NP, MR>;
    using T2 = typename Pair<T>::type;
    const int nx = p.nx, nu = p.nu, nc = p.nc, N = p.N, n = p.n;
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
        const int jj = l - k * nu;  // position of this lane's variable inside block k
        // G_k = C_k psi_k + [0 .. D_k .. 0]   (mpc_qp.py:67,73-78)
        for (int r = 0; r < nc; ++r) {
            T g = T(0);
            if (Ck)
                for (int t = 0; t < nx; ++t) g += Ck[r * nx + t] * ps[t * NP + l];
            if (Dk && jj >= 0 && jj < nu) g += Dk[r * nu + jj];
            Gc[l * L::LDG + k * nc + r] = g;
        }
        // h_k = e_k - C_k (phi_k x0)           (mpc_qp.py:68-72)
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
        // stage cost: P += w_x psi_k' psi_k, q += w_x psi_k'(phi_k x0 - target_k)
        if (p.has_wx) {
            for (int t = 0; t < nx; ++t) {
                const T a = w_x * ps[t * NP + l];
#pragma unroll
                for (int i = 0; i < NP; i += 2) {
                    T2 v = *reinterpret_cast<const T2 *>(ps + t * NP + i);
                    Prow[i] += a * v.x;
                    Prow[i + 1] += a * v.y;
                }
            }
        }
        if (p.q_wx) {
            const T *tg = in[OP_TGT] + k * nx;
            for (int t = 0; t < nx; ++t) qj += (w_x * ps[t * NP + l]) * (xb[t] - tg[t]);
        }
        // psi_{k+1} = A_k psi_k, block column k := B_k ; xbar_{k+1} = A_k xbar_k
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
    // terminal cost: P += w_t psi_N' psi_N, q += w_t psi_N'(phi_N x0 - goal)
    const T *ps = psi + cur * nx * NP;
    const T *xb = xbar + cur * nx;
    if (p.has_wt) {
        for (int t = 0; t < nx; ++t) {
            const T a = w_t * ps[t * NP + l];
#pragma unroll
            for (int i = 0; i < NP; i += 2) {
                T2 v = *reinterpret_cast<const T2 *>(ps + t * NP + i);
                Prow[i] += a * v.x;
                Prow[i + 1] += a * v.y;
            }
        }
    }
    if (p.q_wt) {
        const T *goal = in[OP_GOAL];
        for (int t = 0; t < nx; ++t) qj += (w_t * ps[t * NP + l]) * (xb[t] - goal[t]);
    }
    // + w_u I on the real variables, identity on the padding.
#pragma unroll
    for (int i = 0; i < NP; ++i)
        if (i == l) Prow[i] += (l < n) ? (T)p.w_u : T(1);
    return cur;
}

// ---------------------------------------------------------------------------
// The fused kernel.  MREG: rows of M = G J live in registers (NP <= 16);
// otherwise M overwrites G in shared memory.
// ---------------------------------------------------------------------------
template <typename T, int NP, int MR, bool MREG>  // @phase kernel prologue
__global__ void __launch_bounds__(128) mpc_solve_kernel(const SolveParams p) {
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
    T *ts = qs + 2 * NP;
    T *dfull = qs + 3 * NP;
    T *d2 = qs + 4 * NP;
    T *sc = qs + 5 * NP;
    T *psi = wk + L::oJ;
    T *Jf = wk + L::oJ;

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
    condense_instance<T, NP, MR, false>(p, in, Gc, hs, psi, l, Prow, qj, inst, valid);
    __syncwarp();

    // ---- phase B: Cholesky, row l of L in Prow, columns published in Lc -----  // @phase B cholesky
    T dinv = T(0);
    bool spd = true;
#pragma unroll
    for (int c = 0; c < NP; ++c) {
        const T piv = __shfl_sync(FULL_MASK, Prow[c], c, NP);
        spd = spd && (piv > T(0));
        const T inv = rsqrt_(piv);
        const T lc = Prow[c] * inv;
        Lc[c * L::LDL + l] = lc;
        if (l == c) dinv = inv;
        __syncwarp();
        if (((c + 1) & 1) && c + 1 < NP) Prow[c + 1] -= lc * Lc[c * L::LDL + c + 1];
#pragma unroll
        for (int i = (c + 2) & ~1; i < NP; i += 2) {
            const T2 v = *reinterpret_cast<const T2 *>(Lc + c * L::LDL + i);
            Prow[i] -= lc * v.x;
            Prow[i + 1] -= lc * v.y;
        }
    }
    // J = L^-T, row l per lane: back-substitution over rows from the bottom.  // @phase B J=L^-T
    T Jrow[NP];
#pragma unroll
    for (int c = 0; c < NP; ++c) Jrow[c] = (c == l) ? T(1) : T(0);
    qs[l] = qj;
#pragma unroll
    for (int i = NP - 1; i >= 0; --i) {
        if (l == i) {
#pragma unroll
            for (int c = 0; c < NP; c += 2) {
                T2 v;
                v.x = (c >= i) ? Jrow[c] * dinv : T(0);
                v.y = (c + 1 >= i) ? Jrow[c + 1] * dinv : T(0);
                Jrow[c] = v.x;
                Jrow[c + 1] = v.y;
                *reinterpret_cast<T2 *>(Jf + i * NP + c) = v;
            }
        }
        __syncwarp();
        if (i > 0) {
            const T lik = (l < i) ? Lc[l * L::LDL + i] : T(0);  // L[i][l]
            if ((i & 1)) Jrow[i] -= lik * Jf[i * NP + i];
#pragma unroll
            for (int c = (i + 1) & ~1; c < NP; c += 2) {
                const T2 v = *reinterpret_cast<const T2 *>(Jf + i * NP + c);
                Jrow[c] -= lik * v.x;
                Jrow[c + 1] -= lik * v.y;
            }
        }
    }
    // x = -J J' q  // @phase B x=-P^-1 q
    {
        T tl = T(0);
#pragma unroll
        for (int k = 0; k < NP; ++k) tl += Jf[k * NP + l] * qs[k];
        ts[l] = tl;
    }
    __syncwarp();
    T x = T(0);
#pragma unroll
    for (int c = 0; c < NP; c += 2) {
        const T2 v = *reinterpret_cast<const T2 *>(ts + c);
        x -= Jrow[c] * v.x;
        x -= Jrow[c + 1] * v.y;
    }
    xs[l] = x;
    __syncwarp();

    // M = G J (rows l + s*NP), violations, row norms.  // @phase B M=GJ, violations
    T Mrow[MREG ? MR : 1][NP];
    T viol[MR], gn2[MR], mn2[MR];
    bool rowvalid[MR];
#pragma unroll
    for (int s = 0; s < MR; ++s) {
        const int row = l + s * NP;
        rowvalid[s] = row < m;
        T acc[NP];
#pragma unroll
        for (int c = 0; c < NP; ++c) acc[c] = T(0);
        T vi = T(0), g2 = T(0);
#pragma unroll
        for (int k = 0; k < NP; ++k) {
            const T g = rowvalid[s] ? Gc[k * L::LDG + row] : T(0);
            g2 += g * g;
            vi += g * xs[k];
            if (k & 1) acc[k] += g * Jf[k * NP + k];
#pragma unroll
            for (int c = (k + 1) & ~1; c < NP; c += 2) {
                const T2 v = *reinterpret_cast<const T2 *>(Jf + k * NP + c);
                acc[c] += g * v.x;
                acc[c + 1] += g * v.y;
            }
        }
        T m2 = T(0);
#pragma unroll
        for (int c = 0; c < NP; ++c) m2 += acc[c] * acc[c];
        viol[s] = rowvalid[s] ? vi - hs[row] : T(-1);
        gn2[s] = g2;
        mn2[s] = m2;
        if (MREG) {
#pragma unroll
            for (int c = 0; c < NP; ++c) Mrow[s][c] = acc[c];
        } else {
            // In place: this lane is the only reader of row `row` of G.
#pragma unroll
            for (int c = 0; c < NP; ++c) Gc[c * L::LDG + row] = acc[c];
        }
    }
    auto mget = [&](int s, int c) -> T { return MREG ? Mrow[MREG ? s : 0][c] : Gc[c * L::LDG + l + s * NP]; };
    auto mset = [&](int s, int c, T v) {
        if (MREG)
            Mrow[MREG ? s : 0][c] = v;
        else
            Gc[c * L::LDG + l + s * NP] = v;
    };

    // Tolerance of the violation test, per row: eps * (max(1, |h_i|) + |G_i|).
    T vtol[MR], ginv[MR];
#pragma unroll
    for (int s = 0; s < MR; ++s) {
        const T hi = rowvalid[s] ? abs_(hs[l + s * NP]) : T(1);
        vtol[s] = Num<T>::viol_eps * (fmax(T(1), hi) + sqrt_(gn2[s]));
        ginv[s] = gn2[s] > T(0) ? rsqrt_(gn2[s]) : T(1e30);
    }
    __syncwarp();  // Lc (aliased by Rc) and Jf are dead from here on

    // ---- phase C: dual active-set iteration ---------------------------------  // @phase C select row (step 1)
    const int max_iter = p.max_iter;
    int na = 0, it = 0;
    int st = spd ? 0 : 3;
    bool done = !valid || st != 0 || m == 0;
    bool cont = false;
    int pidx = 0;
    T lam = T(0), lamp = T(0), rinv = T(0);
    int aidx = -1;
    unsigned actbits = 0;
    const T INF = Num<T>::inf();

    while (true) {
        {
            // step 1: most violated inactive row, relative to its norm
            const bool sel = !done && !cont;
            T best = T(0);
            int bi = -1;
#pragma unroll
            for (int s = 0; s < MR; ++s) {
                if (sel && rowvalid[s] && !((actbits >> s) & 1) && viol[s] > vtol[s]) {
                    const T score = viol[s] * ginv[s];
                    if (score > best) {
                        best = score;
                        bi = l + s * NP;
                    }
                }
            }
#pragma unroll
            for (int off = NP / 2; off > 0; off >>= 1) {
                const T ob = __shfl_xor_sync(FULL_MASK, best, off, NP);
                const int obi = __shfl_xor_sync(FULL_MASK, bi, off, NP);
                if (obi >= 0 && (bi < 0 || ob > best || (ob == best && obi < bi))) {
                    best = ob;
                    bi = obi;
                }
            }
            if (sel) {
                if (bi < 0) {
                    done = true;  // primal feasible: optimal
                } else {
                    pidx = bi;
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
        // d = J' n_p = -(row p of M); published by the lane that owns row p.  // @phase C publish d
        const int owner = pidx % NP, pslot = pidx / NP;
        if (act && l == owner) {
            T vp = T(0), m2p = T(0);
#pragma unroll
            for (int s = 0; s < MR; ++s)
                if (s == pslot) {
                    vp = viol[s];
                    m2p = mn2[s];
                }
#pragma unroll
            for (int c = 0; c < NP; ++c) {
                T dc = T(0);
#pragma unroll
                for (int s = 0; s < MR; ++s)
                    if (s == pslot) dc = -mget(s, c);
                dfull[c] = dc;
                d2[c] = (c >= na) ? dc : T(0);
            }
            sc[0] = vp;
            sc[1] = m2p;
        }
        __syncwarp();
        // z = J2 d2 (this lane's component), G z (owned rows), |d2|^2  // @phase C z, Gz
        T z = T(0), a2 = T(0);
        T gz[MR];
#pragma unroll
        for (int s = 0; s < MR; ++s) gz[s] = T(0);
#pragma unroll
        for (int c = 0; c < NP; c += 2) {
            const T2 v = *reinterpret_cast<const T2 *>(d2 + c);
            z += Jrow[c] * v.x;
            z += Jrow[c + 1] * v.y;
            a2 += v.x * v.x;
            a2 += v.y * v.y;
#pragma unroll
            for (int s = 0; s < MR; ++s) {
                gz[s] += mget(s, c) * v.x;
                gz[s] += mget(s, c + 1) * v.y;
            }
        }
        // r = R^-1 d1 (component l on lane l < na)  // @phase C r=R^-1 d
        T rv = (l < na) ? dfull[l] : T(0);
        {
            const int namax = __reduce_max_sync(FULL_MASK, act ? na : 0);
            for (int k = namax - 1; k >= 0; --k) {
                const T rk = __shfl_sync(FULL_MASK, rv * rinv, k, NP);
                if (k < na) {
                    if (l == k)
                        rv = rk;
                    else if (l < k)
                        rv -= Rc[k * L::LDR + l] * rk;
                }
            }
        }
        // step lengths  // @phase C step length, move
        const T cand = (act && l < na && rv > T(0)) ? lam / rv : INF;
        T t1 = cand;
#pragma unroll
        for (int off = NP / 2; off > 0; off >>= 1) t1 = fmin(t1, __shfl_xor_sync(FULL_MASK, t1, off, NP));
        const unsigned bal = __ballot_sync(FULL_MASK, cand == t1 && cand < INF);
        const unsigned segbits = (NP == 32) ? bal : ((bal >> seg_shift) & ((1u << (NP & 31)) - 1u));
        const int lidx = segbits ? (__ffs(segbits) - 1) : 0;
        const T violp = sc[0], dn2 = sc[1];
        const bool zzero = !(a2 > Num<T>::dep_eps * dn2);
        const T t2 = zzero ? INF : violp / a2;
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
            const T alpha = sqrt_(a2);
            const T beta = (dna > T(0)) ? -alpha : alpha;
            const T tau = full ? T(1) / (a2 - beta * dna) : T(0);
            T dj = T(0);
            T dm[MR];
#pragma unroll
            for (int s = 0; s < MR; ++s) dm[s] = T(0);
#pragma unroll
            for (int c = 0; c < NP; c += 2) {
                T2 v = *reinterpret_cast<const T2 *>(d2 + c);
                if (c == na) v.x -= beta;
                if (c + 1 == na) v.y -= beta;
                dj += Jrow[c] * v.x;
                dj += Jrow[c + 1] * v.y;
#pragma unroll
                for (int s = 0; s < MR; ++s) {
                    dm[s] += mget(s, c) * v.x;
                    dm[s] += mget(s, c + 1) * v.y;
                }
            }
            dj *= tau;
#pragma unroll
            for (int s = 0; s < MR; ++s) dm[s] *= tau;
#pragma unroll
            for (int c = 0; c < NP; c += 2) {
                T2 v = *reinterpret_cast<const T2 *>(d2 + c);
                if (c == na) v.x -= beta;
                if (c + 1 == na) v.y -= beta;
                Jrow[c] -= dj * v.x;
                Jrow[c + 1] -= dj * v.y;
#pragma unroll
                for (int s = 0; s < MR; ++s) {
                    mset(s, c, mget(s, c) - dm[s] * v.x);
                    mset(s, c + 1, mget(s, c + 1) - dm[s] * v.y);
                }
            }
            if (full) {
                // new column of R: [d1; beta]
                if (l < na) Rc[na * L::LDR + l] = dfull[l];
                if (l == na) {
                    Rc[na * L::LDR + na] = beta;
                    rinv = T(1) / beta;
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
            if (mover) rinv = T(1) / Rc[l * L::LDR + l];
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
// for the MPCQP host class (qpmpc/mpc_qp.py:28-37).
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
    const int n = p.n, m = p.m, nx = p.nx;

    T *inbase = work + (size_t)ipc * p.inst_stride;
    stage_inputs<T>(p, inbase, inst0, cnt, bar);
    T *wk = work + (size_t)iic * p.inst_stride;
    T *Gc = wk + L::oG;
    T *hs = wk + L::oH;
    T *psi = wk + L::oJ;
    const T *in[OP_COUNT];
#pragma unroll
    for (int o = 0; o < OP_COUNT; ++o) {
        const OperandView &v = p.op[o];
        in[o] = v.ptr ? inbase + v.smem_off + (v.per_instance ? (valid ? iic : 0) * v.sz : 0) : nullptr;
    }
    T Prow[NP];
    T qj;
    const int cur = condense_instance<T, NP, MR, true>(p, in, Gc, hs, psi, l, Prow, qj, inst, valid);
    __syncwarp();
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
    const T *ps = psi + cur * nx * NP;
    if (p.psi_last && l < n)
        for (int t = 0; t < nx; ++t) static_cast<T *>(p.psi_last)[((size_t)inst * nx + t) * n + l] = ps[t * NP + l];
    if (p.phi_last) {
        const T *ph = psi + 2 * nx * NP + 2 * nx + cur * nx * nx;
        for (int t = l; t < nx * nx; t += NP) static_cast<T *>(p.phi_last)[(size_t)inst * nx * nx + t] = ph[t];
    }
}

// ---------------------------------------------------------------------------
// X_{k+1} = A_k X_k + B_k U_k, one thread per (instance, state row) group:
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
