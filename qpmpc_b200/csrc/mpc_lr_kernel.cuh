// mpc_lr_kernel.cuh -- long-horizon kernel for terminal-cost problems: fused condense + dual
// active set that exploits the STRUCTURE of the condensed QP instead of treating it as dense
// (the N = 32 / 64 points of the horizon sweep, BASELINE config 5).
//
// With a terminal cost only (qpmpc/mpc_qp.py:99-105 of the reference: w_x is None) the Hessian is
// the identity plus a matrix of rank nx,
//     P = w_u I + w_t psi_N' psi_N,        psi_N in R^(nx x n),
// so the factor the Goldfarb-Idnani iteration needs -- any J with J J' = P^-1 -- can be written
// down without an O(n^3) Cholesky:
//     J = (I - psi' T psi) / sqrt(w_u),    T = C^-T (I - F) C^-1,
//     K = psi psi' = C C',  F F' = kappa (kappa I + C'C)^-1,  kappa = w_u / w_t
// (3 x 3 algebra; (I - psi'T psi)(I - psi'T psi)' = I - psi'(kappa I + K)^-1 psi = w_u P^-1 by the
// Woodbury identity).  A row of M = G J then costs O(n nx) instead of a length-n triangular
// solve, x = -P^-1 (q + G_A' lambda) is recovered the same way, and no n x n factor is ever
// stored.  G itself is block-Toeplitz for a time-invariant model: one table of n numbers.
//
// Shape handled: time-invariant A, B, C (no D), nu = 1, nc = 2 with paired rows C = [c; -c]
// (desc.paired), w_t set and w_x None, n = N <= NP.  One CTA of NP threads (NP = 32: one warp,
// NP = 64: two warps) owns one instance; thread l is variable l and owns the stored row of step
// l (both signs), kept in REGISTERS as a row of M.  The iteration is the J-free form of the
// warp kernel (mpc_kernels.cuh): rows of M, an explicit R^-1 in shared memory, a Householder
// reflection per added constraint, Givens rotations per dropped one.  Reference citations as
// there: qpmpc/mpc_qp.py:53-105,139-149 (condensing), qpmpc/solve_mpc.py:43 (the QP solve).
#pragma once

#include "mpc_kernels.cuh"  // stage_inputs, Num, rcp_, frsqrt_

namespace qpmpc {

template <typename T, int NP>
struct LrLay {
    static constexpr int NXM = 4;                   // largest nx compiled
    static constexpr int oPsi = 0;                  // psi_N by rows: psi[t*NP + c]
    static constexpr int oGt = oPsi + NXM * NP;     // Toeplitz table of the (+) row: gt[e] = c' A^e B
    static constexpr int oH = oGt + NP;             // h, all 2 N rows
    static constexpr int oRi = oH + 2 * NP;         // R^-1 by columns: Ri[k*NP + row]
    static constexpr int oD = oRi + NP * NP;        // draw, q, d2, tq, cand, lam [NP each]
    static constexpr int oA = oD + 6 * NP;          // aidx [NP] (int), padded to NP elements
    static constexpr int oRed = oA + NP;            // reduction slots: 8 x 8 bytes (16 elements), then 48 values
    static constexpr int oSc = oRed + 16 + 48;      // scalars published with d [8], W = (kappa I + K)^-1 [16]
    static constexpr int fixed = oSc + 8 + 16;      // the staged inputs follow
};

// ---- reductions over the NP threads of the CTA (half a warp, one warp or two) ------------------
// `red` holds two sets of slots used alternately (parity), so one barrier per call is enough.
template <int NP>
struct LrMask {  // the lanes of a warp that exist: a CTA of 16 threads is half a warp
    static constexpr unsigned value = NP >= 32 ? FULL_MASK : ((1u << (NP & 31)) - 1u);
};
template <int NP>
__device__ __forceinline__ unsigned lr_max_u32(unsigned key, unsigned long long *red, int &par) {
    key = __reduce_max_sync(LrMask<NP>::value, key);  // REDUX
    if (NP <= 32) return key;
    unsigned *slot = reinterpret_cast<unsigned *>(red + (par & 1) * 4);
    par ^= 1;
    if ((threadIdx.x & 31) == 0) slot[threadIdx.x >> 5] = key;
    __syncthreads();
    const unsigned a = slot[0], b = slot[1];
    return a > b ? a : b;
}
template <int NP>
__device__ __forceinline__ unsigned long long lr_min_u64(unsigned long long key, unsigned long long *red, int &par) {
    // two REDUX: the high words, then the low words of those that tie
    const unsigned hi = (unsigned)(key >> 32), lo = (unsigned)key;
    const unsigned mh = __reduce_min_sync(LrMask<NP>::value, hi);
    const unsigned ml = __reduce_min_sync(LrMask<NP>::value, hi == mh ? lo : 0xffffffffu);
    key = ((unsigned long long)mh << 32) | ml;
    if (NP <= 32) return key;
    unsigned long long *slot = red + (par & 1) * 4;
    par ^= 1;
    if ((threadIdx.x & 31) == 0) slot[threadIdx.x >> 5] = key;
    __syncthreads();
    const unsigned long long a = slot[0], b = slot[1];
    return a < b ? a : b;
}
template <typename T, int NP, int NV>
__device__ __forceinline__ void lr_sum(T (&v)[NV], T *red, int &par) {
#pragma unroll
    for (int off = (NP >= 32 ? 16 : NP / 2); off > 0; off >>= 1) {
#pragma unroll
        for (int i = 0; i < NV; ++i) v[i] += __shfl_xor_sync(LrMask<NP>::value, v[i], off);
    }
    if (NP <= 32) return;
    static_assert(2 * NV <= 24, "reduction slots");
    T *slot = red + (par & 1) * 24;
    par ^= 1;
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) slot[(threadIdx.x >> 5) * NV + i] = v[i];
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = slot[i] + slot[NV + i];
}
template <int NP>
__device__ __forceinline__ void lr_sync() {
    if (NP <= 32)
        __syncwarp(LrMask<NP>::value);
    else
        __syncthreads();
}

// Cholesky of a symmetric 3 x 3 (or smaller) matrix held in registers, lower factor in place.
template <typename T, int NX>
__device__ __forceinline__ bool chol_small(T (&a)[NX * NX]) {
    bool ok = true;
#pragma unroll
    for (int c = 0; c < NX; ++c) {
        T d = a[c * NX + c];
#pragma unroll
        for (int k = 0; k < c; ++k) d -= a[c * NX + k] * a[c * NX + k];
        ok = ok && (d > T(0));
        const T inv = frsqrt_(d);
        a[c * NX + c] = d * inv;
#pragma unroll
        for (int i = c + 1; i < NX; ++i) {
            T s = a[i * NX + c];
#pragma unroll
            for (int k = 0; k < c; ++k) s -= a[i * NX + k] * a[c * NX + k];
            a[i * NX + c] = s * inv;
        }
#pragma unroll
        for (int i = 0; i < c; ++i) a[i * NX + c] = T(0);
    }
    return ok;
}
// Inverse of a lower-triangular NX x NX matrix (in registers).
template <typename T, int NX>
__device__ __forceinline__ void inv_lower_small(const T (&l)[NX * NX], T (&out)[NX * NX]) {
#pragma unroll
    for (int c = 0; c < NX; ++c) {
#pragma unroll
        for (int i = 0; i < NX; ++i) {
            if (i < c) {
                out[i * NX + c] = T(0);
            } else {
                T s = (i == c) ? T(1) : T(0);
#pragma unroll
                for (int k = c; k < i; ++k) s -= l[i * NX + k] * out[k * NX + c];
                out[i * NX + c] = s * rcp_(l[i * NX + i]);
            }
        }
    }
}

// resident CTAs per SM the NP = 64 variant is compiled for (caps registers at 65536 / (64 * this))
#ifndef QPMPC_LR_MINB
#define QPMPC_LR_MINB 4
#endif
// the same for the one-warp variant (NP = 32): 16 one-warp CTAs per SM at 128 registers -- measured
// 39 -> 62 M solves/s at N = 32 against 8 CTAs at 196 registers (latency-bound: occupancy pays);
// a packed triangular R^-1 (more CTAs by shared memory) was measured slower at equal occupancy
#ifndef QPMPC_LR_MINB32
#define QPMPC_LR_MINB32 16
#endif
// NP = 16: a CTA is HALF a warp (16 threads) -- lanes idle, but no second instance in lockstep,
// no CTA tail and every branch uniform; a row of M is 32 registers, so more CTAs fit
#ifndef QPMPC_LR_MINB16
#define QPMPC_LR_MINB16 24
#endif

template <typename T, int NP, int NX>  // @phase LR kernel
__global__ void __launch_bounds__(NP, NP == 16 ? QPMPC_LR_MINB16 : NP == 32 ? QPMPC_LR_MINB32 : QPMPC_LR_MINB) mpc_solve_lr_kernel(const SolveParams p) {
    using L = LrLay<T, NP>;
    using T2 = typename Pair<T>::type;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw);
    T *sm = reinterpret_cast<T *>(smem_raw + 16);
    T *psis = sm + L::oPsi, *gt = sm + L::oGt, *hs = sm + L::oH, *Ri = sm + L::oRi;
    T *draw = sm + L::oD, *qs = draw + NP, *d2 = draw + 2 * NP, *tq = draw + 3 * NP, *cands = draw + 4 * NP;
    T *lams = draw + 5 * NP;
    int *aidxs = reinterpret_cast<int *>(sm + L::oA);
    unsigned long long *redk = reinterpret_cast<unsigned long long *>(sm + L::oRed);
    T *redv = sm + L::oRed + 16;
    T *sc = sm + L::oSc;
    T *inbase = sm + L::fixed;

    const int l = threadIdx.x;
    const long long inst = blockIdx.x;
    const int N = p.N, n = p.n, m = p.m;
    int par = 0;

    stage_inputs<T>(p, inbase, (int)inst, 1, bar);
    const T *in[OP_COUNT];
#pragma unroll
    for (int o = 0; o < OP_COUNT; ++o) in[o] = p.op[o].ptr ? inbase + p.op[o].smem_off : nullptr;

    // ---- phase A: condensing (qpmpc/mpc_qp.py:53-105,139-149) -------------------------  // @phase LR condense
    // psi_N[:, l] = A^(N-1-l) B and y_l = A^l x0 from the binary digits of the exponents while A
    // is squared log2(N) times; thread k < N owns the h rows of step k; x_N = A y_(N-1).
    T psi[NX], xN[NX], qj;
    {
        T Ar[NX * NX], Ap[NX * NX], y[NX], c0[NX], c1[NX];
#pragma unroll
        for (int t = 0; t < NX * NX; ++t) Ar[t] = Ap[t] = in[OP_A][t];
#pragma unroll
        for (int t = 0; t < NX; ++t) {
            psi[t] = (l < n) ? in[OP_B][t] : T(0);
            y[t] = in[OP_X0][t];
            c0[t] = in[OP_C][t];
            c1[t] = in[OP_C][NX + t];
        }
        const int ev = (l < n) ? N - 1 - l : 0;
#pragma unroll 1
        for (int bit = 1; bit < N; bit <<= 1) {
            const bool sv = (ev & bit) != 0, sy = (l & bit) != 0;
            T nv[NX], ny[NX];
#pragma unroll
            for (int t = 0; t < NX; ++t) {
                T a = T(0), b = T(0);
#pragma unroll
                for (int s = 0; s < NX; ++s) {
                    a += Ap[t * NX + s] * psi[s];
                    b += Ap[t * NX + s] * y[s];
                }
                nv[t] = a;
                ny[t] = b;
            }
#pragma unroll
            for (int t = 0; t < NX; ++t) {
                psi[t] = sv ? nv[t] : psi[t];
                y[t] = sy ? ny[t] : y[t];
            }
            if ((bit << 1) < N) {
                T sq[NX * NX];
#pragma unroll
                for (int t = 0; t < NX * NX; ++t) {
                    T a = T(0);
#pragma unroll
                    for (int s = 0; s < NX; ++s) a += Ap[(t / NX) * NX + s] * Ap[s * NX + t % NX];
                    sq[t] = a;
                }
#pragma unroll
                for (int t = 0; t < NX * NX; ++t) Ap[t] = sq[t];
            }
        }
        if (l < N) {
            const T *ek = in[OP_E] + l * p.op[OP_E].step;
            T h0 = ek[0], h1 = ek[1];
#pragma unroll
            for (int t = 0; t < NX; ++t) {
                h0 -= c0[t] * y[t];
                h1 -= c1[t] * y[t];
            }
            hs[2 * l] = h0;
            hs[2 * l + 1] = h1;
        }
        // psi rows, the Toeplitz table of the (+) row, y_(N-1)
        T g = T(0);
#pragma unroll
        for (int t = 0; t < NX; ++t) {
            psis[t * NP + l] = psi[t];
            g += c0[t] * psi[t];
        }
        if (l < n) gt[n - 1 - l] = g;
        if (l == N - 1) {
#pragma unroll
            for (int t = 0; t < NX; ++t) redv[t] = y[t];
        }
        lr_sync<NP>();
#pragma unroll
        for (int t = 0; t < NX; ++t) {
            T b = T(0);
#pragma unroll
            for (int s = 0; s < NX; ++s) b += Ar[t * NX + s] * redv[s];
            xN[t] = b;
        }
        qj = T(0);
        if (p.q_wt) {
#pragma unroll
            for (int t = 0; t < NX; ++t) qj += ((T)p.w_t * psi[t]) * (xN[t] - in[OP_GOAL][t]);
        }
    }

    // ---- the 3 x 3 algebra behind J and P^-1 (every thread, redundantly)  // @phase LR small algebra
    // K = psi psi' (sum over the threads), T for J, W = (kappa I + K)^-1 for P^-1.
    T Tm[NX * NX], Wm[NX * NX];
    bool spd = true;
    const T w_u = (T)p.w_u, kappa = (T)(p.w_u / p.w_t), rsw = frsqrt_(w_u);
    {
        constexpr int NK = NX * (NX + 1) / 2;
        T kv[NK];
        {
            int idx = 0;
#pragma unroll
            for (int i = 0; i < NX; ++i)
#pragma unroll
                for (int j = 0; j <= i; ++j) kv[idx++] = psi[i] * psi[j];
        }
        lr_sync<NP>();  // redv was read above
        lr_sum<T, NP, NK>(kv, redv, par);
        T K[NX * NX], C[NX * NX], Ci[NX * NX], E[NX * NX], F[NX * NX];
        {
            int idx = 0;
#pragma unroll
            for (int i = 0; i < NX; ++i)
#pragma unroll
                for (int j = 0; j <= i; ++j) {
                    K[i * NX + j] = K[j * NX + i] = kv[idx];
                    ++idx;
                }
        }
#pragma unroll
        for (int t = 0; t < NX * NX; ++t) C[t] = K[t];
        spd = chol_small<T, NX>(C);
        inv_lower_small<T, NX>(C, Ci);
        // E = kappa I + C'C, then E^-1 = Li' Li with Li = chol(E)^-1; F = chol(kappa E^-1)
#pragma unroll
        for (int i = 0; i < NX; ++i)
#pragma unroll
            for (int j = 0; j < NX; ++j) {
                T s = (i == j) ? kappa : T(0);
#pragma unroll
                for (int k = 0; k < NX; ++k) s += C[k * NX + i] * C[k * NX + j];
                E[i * NX + j] = s;
            }
        T Le[NX * NX], Lei[NX * NX];
#pragma unroll
        for (int t = 0; t < NX * NX; ++t) Le[t] = E[t];
        spd = chol_small<T, NX>(Le) && spd;
        inv_lower_small<T, NX>(Le, Lei);
#pragma unroll
        for (int i = 0; i < NX; ++i)
#pragma unroll
            for (int j = 0; j < NX; ++j) {
                T s = T(0);
#pragma unroll
                for (int k = 0; k < NX; ++k) s += Lei[k * NX + i] * Lei[k * NX + j];
                F[i * NX + j] = kappa * s;  // kappa E^-1
            }
        spd = chol_small<T, NX>(F) && spd;
        // T = Ci' (I - F) Ci
        T Y[NX * NX], YC[NX * NX];
#pragma unroll
        for (int i = 0; i < NX; ++i)
#pragma unroll
            for (int j = 0; j < NX; ++j) Y[i * NX + j] = ((i == j) ? T(1) : T(0)) - F[i * NX + j];
#pragma unroll
        for (int i = 0; i < NX; ++i)
#pragma unroll
            for (int j = 0; j < NX; ++j) {
                T s = T(0);
#pragma unroll
                for (int k = 0; k < NX; ++k) s += Y[i * NX + k] * Ci[k * NX + j];
                YC[i * NX + j] = s;
            }
#pragma unroll
        for (int i = 0; i < NX; ++i)
#pragma unroll
            for (int j = 0; j < NX; ++j) {
                T s = T(0);
#pragma unroll
                for (int k = 0; k < NX; ++k) s += Ci[k * NX + i] * YC[k * NX + j];
                Tm[i * NX + j] = s;
            }
        // W = (kappa I + K)^-1 = Lw' Lw, Lw = chol(kappa I + K)^-1
        T Kk[NX * NX], Lwi[NX * NX];
#pragma unroll
        for (int t = 0; t < NX * NX; ++t) Kk[t] = K[t] + ((t / NX == t % NX) ? kappa : T(0));
        spd = chol_small<T, NX>(Kk) && spd;
        inv_lower_small<T, NX>(Kk, Lwi);
#pragma unroll
        for (int i = 0; i < NX; ++i)
#pragma unroll
            for (int j = 0; j < NX; ++j) {
                T s = T(0);
#pragma unroll
                for (int k = 0; k < NX; ++k) s += Lwi[k * NX + i] * Lwi[k * NX + j];
                Wm[i * NX + j] = s;
            }
    }

    // (kept for phase D in shared memory, not in registers, like psi)
    qs[l] = qj;
    if (l == 0) {
#pragma unroll
        for (int t = 0; t < NX * NX; ++t) sc[8 + t] = Wm[t];
    }

    // ---- phase B: t = J'q, the owned row of M = G J, violations  // @phase LR rows of M
    T Mrow[NP];
    T viol, cen, wid, vtol, ginv, mn2;
    const bool rowvalid = l < N;  // stored row l = the pair of step l
    {
        // t = J' q = (q - psi' T' (psi q)) / sqrt(w_u)
        T pq[NX];
#pragma unroll
        for (int t = 0; t < NX; ++t) pq[t] = psi[t] * qj;
        lr_sum<T, NP, NX>(pq, redv, par);
        T tl = qj;
#pragma unroll
        for (int i = 0; i < NX; ++i) {
            T s = T(0);
#pragma unroll
            for (int j = 0; j < NX; ++j) s += Tm[j * NX + i] * pq[j];  // (T' pq)_i
            tl -= psi[i] * s;
        }
        tq[l] = tl * rsw;
        // G row of step l: G[c] = gt[l - 1 - c] for c < l, 0 beyond; a = G psi'
        T a[NX], g2 = T(0);
#pragma unroll
        for (int t = 0; t < NX; ++t) a[t] = T(0);
#pragma unroll
        for (int c = 0; c < NP; ++c) {
            const T g = (rowvalid && c < l) ? gt[l - 1 - c] : T(0);
            Mrow[c] = g;
            g2 += g * g;
#pragma unroll
            for (int t = 0; t < NX; ++t) a[t] += g * psis[t * NP + c];
        }
        T b[NX];
#pragma unroll
        for (int j = 0; j < NX; ++j) {
            T s = T(0);
#pragma unroll
            for (int i = 0; i < NX; ++i) s += a[i] * Tm[i * NX + j];
            b[j] = s;
        }
        lr_sync<NP>();  // tq complete
        T m2 = T(0), mt = T(0);
#pragma unroll
        for (int c = 0; c < NP; ++c) {
            T v = Mrow[c];
#pragma unroll
            for (int t = 0; t < NX; ++t) v -= b[t] * psis[t * NP + c];
            v *= rsw;
            Mrow[c] = v;
            m2 += v * v;
            mt += v * tq[c];
        }
        const T hp = rowvalid ? hs[2 * l] : T(0), hm = rowvalid ? hs[2 * l + 1] : T(0);
        cen = rowvalid ? T(0.5) * (hp - hm) : T(0);
        wid = rowvalid ? T(0.5) * (hp + hm) : T(1);
        viol = rowvalid ? -mt : T(0);  // G x for x = -P^-1 q = -J t
        vtol = Num<T>::viol_eps * (fmax(T(1), fmax(abs_(hp), abs_(hm))) + sqrt_(g2));
        ginv = g2 > T(0) ? frsqrt_(g2) : T(1e30);
        mn2 = m2;
    }
    // R^-1 starts as the zero matrix (entries below the diagonal stay zero throughout)
#pragma unroll 4
    for (int k = 0; k < NP; ++k) Ri[k * NP + l] = T(0);
    lams[l] = T(0);
    aidxs[l] = -1;

    // ---- phase C: dual active-set iteration (one instance per CTA: every branch is uniform)  // @phase LR loop
    int na = 0, it = 0, st = spd ? 0 : 3, pidx = 0;
    bool pneg = false, cont = false, active = false;  // active: this thread's pair is in the working set
    T lamp = T(0);
    const T INF = Num<T>::inf();
    {
        // a pair with h+ + h- < 0 admits no point at all
        const unsigned bad = lr_max_u32<NP>((rowvalid && wid < -vtol) ? 1u : 0u, redk, par);
        if (bad && st == 0) st = 2;
    }
    lr_sync<NP>();
    while (st == 0 && m > 0) {
        if (!cont) {
            const T off = viol - cen;
            const T vs = abs_(off) - wid;
            const float score = (float)(vs * ginv);
            // single-precision key with the row and its sign in the low 7 bits (the ranking is a
            // heuristic: any violated row is a valid pivot; ties go to the lowest row)
            unsigned key = 0u;
            if (rowvalid && !active && vs > vtol && score > 0.f)
                key = (__float_as_uint(fmaxf(score, 1e-37f)) & ~127u) | (unsigned)((NP - 1 - l) << 1) |
                      (off < T(0) ? 1u : 0u);
            key = lr_max_u32<NP>(key, redk, par);
            if (key == 0u) break;  // primal feasible: optimal
            pidx = NP - 1 - (int)((key & 127u) >> 1);
            pneg = (key & 1u) != 0;
            lamp = T(0);
        }
        if (++it > p.max_iter) {
            st = 1;
            break;
        }
        // row p of M (with its sign) is -d
        if (l == pidx) {
            const T sg = pneg ? T(-1) : T(1);
#pragma unroll
            for (int c = 0; c < NP; c += 2) {
                T2 v;
                v.x = sg * Mrow[c];
                v.y = sg * Mrow[c + 1];
                *reinterpret_cast<T2 *>(draw + c) = v;
            }
            sc[0] = abs_(viol - cen) - wid;
            sc[1] = mn2;
        }
        lr_sync<NP>();
        const T dl = draw[l];
        const T d2l = (l >= na) ? dl : T(0);
        d2[l] = d2l;  // the part of d outside the working set, zero-padded; the rest is read from draw
        T a2v[1] = {d2l * d2l};
        lr_sum<T, NP, 1>(a2v, redv, par);  // (its barrier also publishes d2)
        if (NP <= 32) __syncwarp(LrMask<NP>::value);
        const T a2 = a2v[0];
        // -G z = M2 m2 for the owned row
        T gz0 = T(0), gz1 = T(0), gz2 = T(0), gz3 = T(0);  // four chains: the kernel is latency-bound
#pragma unroll
        for (int c = 0; c < NP; c += 4) {
            const T2 v = *reinterpret_cast<const T2 *>(d2 + c);
            const T2 w = *reinterpret_cast<const T2 *>(d2 + c + 2);
            gz0 += Mrow[c] * v.x;
            gz1 += Mrow[c + 1] * v.y;
            gz2 += Mrow[c + 2] * w.x;
            gz3 += Mrow[c + 3] * w.y;
        }
        const T gz = (gz0 + gz1) + (gz2 + gz3);
        // -r = R^-1 m1 (component l; zero on threads >= na)
        T rv = T(0);
        {
            T rv1 = T(0);
            for (int k = 0; k + 1 < na; k += 2) {
                const T2 dk = *reinterpret_cast<const T2 *>(draw + k);
                rv += Ri[k * NP + l] * dk.x;
                rv1 += Ri[(k + 1) * NP + l] * dk.y;
            }
            if (na & 1) rv += Ri[(na - 1) * NP + l] * draw[na - 1];
            rv += rv1;
        }
        const T lam = lams[l];
        const T cand = (l < na && rv < T(0)) ? fmax(lam, T(0)) * rcp_(-rv) : INF;
        cands[l] = cand;
        // smallest ratio, ties to the lowest position: the low 6 bits of the key carry the position
        unsigned long long kmin = ((unsigned long long)__double_as_longlong((double)cand) & ~63ull) | (unsigned)l;
        kmin = lr_min_u64<NP>(kmin, redk, par);  // (two warps: its barrier also publishes cands)
        if (NP <= 32) __syncwarp(LrMask<NP>::value);
        const int lidx = (int)(kmin & 63ull);
        const T t1 = cands[lidx];
        const T violp = sc[0], dn2 = sc[1];
        const bool zzero = !(a2 > Num<T>::dep_eps * dn2);
        const T ainv = frsqrt_(a2);
        const T t2 = zzero ? INF : violp * (ainv * ainv);
        if (t1 == INF && t2 == INF) {
            st = 2;  // infeasible
            break;
        }
        const T t = t2 < t1 ? t2 : t1;
        if (!zzero) viol -= t * gz;
        if (l < na) lams[l] = lam + t * rv;
        lamp += t;
        const bool full = !zzero && t2 <= t1;
        if (full) {  // @phase LR add constraint
            // Householder: reflect d2 onto beta e_na, applied to the columns >= na of M
            const T mna = draw[na < NP ? na : NP - 1];
            const T alpha = a2 * ainv;
            const T beta = (mna < T(0)) ? -alpha : alpha;
            const T binv = (mna < T(0)) ? -ainv : ainv;
            const T tau = rcp_(a2 + beta * mna);
            // -v = d2 + beta e_na: M v is the product with d2 computed above plus one entry of the
            // row (na is uniform: a switch over registers), and the update is one pass over d2
            const T dm = (gz + reg_get<T, NP>(Mrow, na) * beta) * tau;
#pragma unroll
            for (int c = 0; c < NP; c += 2) {
                const T2 v = *reinterpret_cast<const T2 *>(d2 + c);
                Mrow[c] -= dm * v.x;
                Mrow[c + 1] -= dm * v.y;
            }
            reg_sub<T, NP>(Mrow, na, dm * beta);
            // R gains the column [d1; beta]: R^-1 gains [-r / beta; 1 / beta]
            if (l < na) Ri[na * NP + l] = rv * binv;
            if (l == na) {
                Ri[na * NP + na] = binv;
                lams[na] = lamp;
                aidxs[na] = pidx | (pneg ? 0x8000 : 0);
            }
            if (l == pidx) active = true;
            ++na;
            cont = false;
            lr_sync<NP>();
        } else {  // @phase LR drop constraint
            // the constraint at active position lidx leaves; p stays the candidate
            const int nan_ = na - 1;
            const int cidx = aidxs[lidx] & 0x7fff;
            if (l == cidx) active = false;
            // row lidx of R^-1 fixes the rotations of adjacent columns (j, j+1), j = lidx .. na-2,
            // that zero it up to its last entry; they act on thread-private rows of R^-1 and M
            T *rrow = tq;  // (free since phase B)
            if (l < na) rrow[l] = Ri[l * NP + lidx];  // row lidx, by columns
            lr_sync<NP>();
            const T lam_n = (l + 1 < na) ? lams[l + 1] : T(0);
            const int aidx_n = (l + 1 < na) ? aidxs[l + 1] : -1;
            T a = rrow[lidx];
#pragma unroll
            for (int j = 0; j < NP - 1; ++j) {
                if (j < lidx || j >= nan_) continue;
                const T b = rrow[j + 1];
                const T h2 = a * a + b * b;
                const T hinv = h2 > T(0) ? frsqrt_(h2) : T(0);
                const T cs = h2 > T(0) ? b * hinv : T(1);
                const T sn = -a * hinv;
                a = h2 * hinv;
                if (l <= j + 1) {
                    const T u = (l <= j) ? Ri[j * NP + l] : T(0), v = Ri[(j + 1) * NP + l];
                    Ri[j * NP + l] = cs * u + sn * v;
                    Ri[(j + 1) * NP + l] = cs * v - sn * u;
                }
                const T mu = Mrow[j], mv = Mrow[j + 1];
                Mrow[j] = cs * mu + sn * mv;
                Mrow[j + 1] = cs * mv - sn * mu;
            }
            lr_sync<NP>();
            if (l >= lidx && l < nan_) {
                lams[l] = lam_n;
                aidxs[l] = aidx_n;
            }
            if (l == nan_) {
                lams[l] = T(0);
                aidxs[l] = -1;
            }
            // remove row lidx: the rows above it move down one position
            for (int k = 0; k < nan_; ++k) {
                const bool mv = l >= lidx && l < nan_ && l <= k;
                T v = T(0);
                if (mv) v = Ri[k * NP + l + 1];
                lr_sync<NP>();
                if (mv) Ri[k * NP + l] = v;
                if (l == k + 1 && l > lidx) Ri[k * NP + l] = T(0);
            }
            na = nan_;
            cont = true;
            lr_sync<NP>();
        }
    }

    // ---- phase D: x = -P^-1 (q + G_A' lambda), P^-1 = (I - psi' W psi) / w_u  // @phase LR outputs
    lr_sync<NP>();
    T x = T(0);
    {
        T w = qs[l];
        T psl[NX];
#pragma unroll
        for (int t = 0; t < NX; ++t) psl[t] = psis[t * NP + l];
        if (st == 0) {
            for (int i = 0; i < na; ++i) {
                const int ai = aidxs[i];
                const int k = ai & 0x7fff;  // step of the pair; its (+) row is G[c] = gt[k-1-c], c < k
                const T li = (ai & 0x8000) ? -lams[i] : lams[i];
                if (l < k) w += li * gt[k - 1 - l];
            }
        }
        T pw[NX];
#pragma unroll
        for (int t = 0; t < NX; ++t) pw[t] = psl[t] * w;
        lr_sync<NP>();
        lr_sum<T, NP, NX>(pw, redv, par);
        T corr = T(0);
#pragma unroll
        for (int i = 0; i < NX; ++i) {
            T s = T(0);
#pragma unroll
            for (int j = 0; j < NX; ++j) s += sc[8 + i * NX + j] * pw[j];
            corr += psl[i] * s;
        }
        x = -(w - corr) * rcp_(w_u);
    }
    {
        const unsigned bad = lr_max_u32<NP>((l < n && !(abs_(x) < INF)) ? 1u : 0u, redk, par);
        if (st == 0 && bad) st = 3;
    }
    const T xo = (st == 0) ? x : Num<T>::nan();
    if (l < n) {
        if (p.U) static_cast<T *>(p.U)[(size_t)inst * n + l] = xo;
        for (int r = 0; r < p.npeers; ++r) static_cast<T *>(p.peerU[r])[(size_t)(p.row_off + inst) * n + l] = xo;
    }
    if (l == 0) {
        if (p.status) p.status[inst] = st;
        if (p.iters) p.iters[inst] = it;
        for (int r = 0; r < p.npeers; ++r)
            if (p.peer_status[r]) p.peer_status[r][p.row_off + inst] = st;
    }
    if (p.Z) {
        T *Zb = static_cast<T *>(p.Z) + (size_t)inst * m;
        for (int r = l; r < m; r += NP) Zb[r] = T(0);
        lr_sync<NP>();
        if (st == 0 && l < na) {
            const int ai = aidxs[l];
            Zb[(ai & 0x7fff) * 2 + ((ai & 0x8000) ? 1 : 0)] = lams[l];
        }
    }
}

// Shared-memory bytes of one CTA, and where the staged operands go.
template <typename T, int NP>
size_t lr_layout_smem(SolveParams *p) {
    using L = LrLay<T, NP>;
    int off = 0;
    p->present_mask = 0;
    for (int o = 0; o < OP_COUNT; ++o) {
        OperandView &v = p->op[o];
        if (!v.ptr) continue;
        p->present_mask |= 1 << o;
        v.smem_off = off;
        off += (v.sz + 3) / 4 * 4;
    }
    p->input_elems = off;
    return 16 + ((size_t)L::fixed + off) * sizeof(T);
}

// Whether horizons 8 < n <= 16 take this kernel (half-warp CTAs) or the warp kernel
// (QPMPC_B200_LR16 overrides).
constexpr int LR16_DEFAULT = 0;

// The shape the kernel handles (see the header of this file).
inline bool lr_applicable(const SolveParams &p, bool paired) {
    const bool lti = p.op[OP_A].step == 0 && p.op[OP_B].step == 0 && p.op[OP_C].ptr && p.op[OP_C].step == 0;
    return paired && lti && !p.op[OP_D].ptr && p.nu == 1 && p.nc == 2 && p.has_wt && !p.has_wx && p.w_t > 0.0 &&
           (p.nx == 2 || p.nx == 3 || p.nx == 4) && p.n <= 64;
}

}  // namespace qpmpc
