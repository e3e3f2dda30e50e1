// mpc_pdip.cuh -- fused condense + primal-dual interior-point QP kernel
// (sm_100a): the method BASELINE.json's north star names, behind
// desc.method = QPMPC_B200_PDIP.  The default method stays the exact dual
// active set of mpc_kernels.cuh (DESIGN.md section 2 says why); this kernel
// trades exactness for an iteration count that does not depend on the number
// of active constraints.
//
// Same ownership as the active-set kernel: a group of NP lanes owns one MPC
// instance, lane l is decision variable l and owns the constraint rows
// l, l + NP, ... (MR of them); a warp carries 32 / NP instances in lock step.
//
//   0  operands staged into shared memory by bulk TMA (stage_inputs)
//   A  condensing (condense_dispatch, qpmpc/mpc_qp.py:53-105,139-149 of the
//      reference): G (dense, by columns), h in shared memory, row l of P and
//      q_l in registers; P is then parked in shared memory
//   B  Mehrotra predictor-corrector on   min 1/2 u'Pu + q'u  s.t. Gu + s = h,
//      s, z >= 0 -- what qpsolvers.solve_problem (qpmpc/solve_mpc.py:43) does
//      in an interior-point backend.  Per iteration: residuals, the reduced
//      KKT matrix H = P + G' diag(z/s) G (row l per lane, in registers), its
//      Cholesky factor (columns published in shared memory), and two
//      forward/backward substitutions by warp shuffles
//   C  primal-dual active-set polish: rows with z > s form the active set A; a
//      few proximal multiplier steps in residual form on the equality QP of A
//      reuse the same build / factor / solve code with W = 1_A / delta; a
//      point that fails the KKT test corrects A by one row and repeats (<= 8 rounds).
//      An accepted point is the exact solution, as the active-set kernel
//      returns it: this is what makes |dU| <= 1e-6 hold at w_u = 1e-6
//   D  U (coalesced), status, iteration count, multipliers
//
// oracle/pdip_np.py is the NumPy statement of phases B and C, line for line.
//
// The same source also compiles for the host: tests/emu/ runs it thread by
// thread on fibers (QPMPC_HOST_EMU) and checks it against the NumPy model and
// the exact oracle without a GPU.
#pragma once

#include "mpc_kernels.cuh"  // stage_inputs, condense_dispatch, fsolve

namespace qpmpc {

template <typename T> struct PdipNum;
template <> struct PdipNum<double> {
    static constexpr double tol_min = 1e-13;    // below this the residuals are rounding noise
    static constexpr double delta = 1e-7;       // proximal parameter of the polish
    static constexpr double polish_eps = 1e-9;  // acceptance test of the polished point
    static constexpr double stat_eps = 1e-9;    // stationarity: |r| <= stat_eps lambda_min(P) |u|, i.e. |du| <= 1e-9 |u|
    static constexpr double macheps = 2.3e-16;
    static constexpr double step_frac = 0.99;
    static constexpr int polish_rounds = 8;
    static constexpr int polish_steps = 8;      // proximal steps per round, at most
};
template <> struct PdipNum<float> {
    static constexpr float tol_min = 1e-6f;
    static constexpr float delta = 1e-3f;
    static constexpr float polish_eps = 1e-4f;
    static constexpr float stat_eps = 1e-3f;
    static constexpr float macheps = 1.2e-7f;
    static constexpr float step_frac = 0.99f;
    static constexpr int polish_rounds = 8;
    static constexpr int polish_steps = 8;
};

// Per-instance shared-memory regions of the interior-point kernel (elements of
// T); the dense G and the scratch of the generic condensing follow at
// tail_layout(fixed, szG, ..., toeplitz = false, mreg = false).
template <typename T, int NP, int MR>
struct PdipLay {
    static constexpr int MP = MR * NP;    // padded constraint rows
    static constexpr int LDG = MP + 1;    // G by columns: Gc[c*LDG + row] (as Lay::LDG)
    static constexpr int LDL = NP + 2;    // L, P by columns
    static constexpr int szG = ((NP * LDG + 3) / 4) * 4;
    static constexpr int szL = ((NP * LDL + 3) / 4) * 4;
    static constexpr int oH = 0;          // hs[MP]
    static constexpr int oL = oH + MP;    // Lc (psi exchange buffers during phase A)
    static constexpr int oP = oL + szL;   // Pc: P by columns (symmetric: column l is row l)
    static constexpr int oV = oP + szL;   // xs[NP], dv[NP]
    static constexpr int oW = oV + 2 * NP;  // wv[MP], tv[MP]
    static constexpr int fixed = oW + 2 * MP;
    static_assert(8 * NP <= szL, "psi exchange buffers must fit in the L region");
};

template <typename T, int NP>
__device__ __forceinline__ T pdip_sum(T v) {
#pragma unroll
    for (int off = NP / 2; off > 0; off >>= 1) v += __shfl_xor_sync(FULL_MASK, v, off, NP);
    return v;
}
template <typename T, int NP>
__device__ __forceinline__ T pdip_min(T v) {
#pragma unroll
    for (int off = NP / 2; off > 0; off >>= 1) v = fmin(v, __shfl_xor_sync(FULL_MASK, v, off, NP));
    return v;
}
template <typename T, int NP>
__device__ __forceinline__ T pdip_max(T v) {
#pragma unroll
    for (int off = NP / 2; off > 0; off >>= 1) v = fmax(v, __shfl_xor_sync(FULL_MASK, v, off, NP));
    return v;
}

// Cholesky H = L L' with row l of H in this lane's registers (destroyed).
// Column c of L is published as Lc[c*LDL + row], 1 / L_cc as dv[c].  Returns
// false (to every lane of the group) if a pivot is not positive.
template <typename T, int NP, int LDL>
__device__ __forceinline__ bool pdip_cholesky(T (&Hrow)[NP], T *Lc, T *dv, int l) {
    bool spd = true;
#pragma unroll
    for (int c = 0; c < NP; ++c) {
        const T piv = __shfl_sync(FULL_MASK, Hrow[c], c, NP);
        spd = spd && (piv > T(0));
        const T inv = frsqrt_(piv);
        const T lc = Hrow[c] * inv;
        Lc[c * LDL + l] = lc;
        if (l == c) dv[c] = inv;
        __syncwarp();
#pragma unroll
        for (int i = c + 1; i < NP; ++i) Hrow[i] -= lc * Lc[c * LDL + i];
    }
    return spd;
}

// Component l of the solution of L L' y = w (w_l per lane): lane c finalises
// component c and broadcasts it, forwards then backwards.
template <typename T, int NP, int LDL>
__device__ __forceinline__ T pdip_solve(const T *Lc, const T *dv, T w, int l) {
    const T dinv = dv[l];
    T y = T(0);
#pragma unroll
    for (int c = 0; c < NP; ++c) {
        const T yc = __shfl_sync(FULL_MASK, w * dinv, c, NP);
        if (l == c) y = yc;
        if (l > c) w -= Lc[c * LDL + l] * yc;
    }
    T s2 = y, x = T(0);
#pragma unroll
    for (int c = NP - 1; c >= 0; --c) {
        const T xc = __shfl_sync(FULL_MASK, s2 * dinv, c, NP);
        if (l == c) x = xc;
        if (l < c) s2 -= Lc[l * LDL + c] * xc;
    }
    return x;
}

// LS variant of the two substitutions: L is replaced once per factorisation by
// its inverse (column l of L^-1 is a lane-local forward substitution with
// L y = e_l, as the active-set kernel builds J), after which a solve is two
// matrix-vector products through shared memory -- 4 __syncwarp instead of 2 NP
// dependent shuffles.  On exit Lc holds L^-1 by columns: Lc[k*LDL + r] = Linv[r][k].
template <typename T, int NP, int LDL>
__device__ __forceinline__ void pdip_invert(T *Lc, const T *dv, int l) {
    T jr[NP];
#pragma unroll
    for (int c = 0; c < NP; ++c) jr[c] = (c == l) ? T(1) : T(0);
    fsolve<T, NP, LDL, 1>(Lc, dv, jr, jr, jr);
    __syncwarp();  // every lane has read L
#pragma unroll
    for (int c = 0; c < NP; ++c) Lc[l * LDL + c] = jr[c];
    __syncwarp();
}
// Component l of H^-1 w = L^-T (L^-1 w); xs [NP] is the exchange vector (free on entry and on exit).
template <typename T, int NP, int LDL>
__device__ __forceinline__ T pdip_solve_inv(const T *Li, T *xs, T w, int l) {
    xs[l] = w;
    __syncwarp();
    T y0 = T(0), y1 = T(0);
#pragma unroll
    for (int k = 0; k < NP; k += 2) {
        y0 += Li[k * LDL + l] * xs[k];
        y1 += Li[(k + 1) * LDL + l] * xs[k + 1];
    }
    __syncwarp();
    xs[l] = y0 + y1;
    __syncwarp();
    T x0 = T(0), x1 = T(0);
#pragma unroll
    for (int c = 0; c < NP; c += 2) {
        x0 += Li[l * LDL + c] * xs[c];
        x1 += Li[l * LDL + c + 1] * xs[c + 1];
    }
    __syncwarp();
    return x0 + x1;
}

// Row l of H = P + G' diag(wv) G into Hrow and, in the same pass over the
// rows, g = sum_row G[row, l] tv[row].  (Skipping the structurally zero columns of the block
// lower triangular G, four at a time behind a warp-uniform branch, was measured: 14.4 -> 12.7
// M solves/s on config 2 -- the straight unrolled stream of FMAs wins.)
template <typename T, int NP, int LDG, int LDL>
__device__ __forceinline__ T pdip_build(T (&Hrow)[NP], const T *Pc, const T *Gc, const T *wv, const T *tv, int m,
                                        int l) {
#pragma unroll
    for (int c = 0; c < NP; ++c) Hrow[c] = Pc[c * LDL + l];
    T g = T(0);
    const T *gl = Gc + l * LDG;
    for (int row = 0; row < m; ++row) {
        const T glr = gl[row];
        const T a = glr * wv[row];
        g += glr * tv[row];
#pragma unroll
        for (int c = 0; c < NP; ++c) Hrow[c] += a * Gc[c * LDG + row];
    }
    return g;
}

// sum_row G[row, l] v[row]
template <typename T, int LDG>
__device__ __forceinline__ T pdip_gt_dot(const T *Gc, const T *v, int m, int l) {
    const T *gl = Gc + l * LDG;
    T g0 = T(0), g1 = T(0);
    int row = 0;
    for (; row + 1 < m; row += 2) {
        g0 += gl[row] * v[row];
        g1 += gl[row + 1] * v[row + 1];
    }
    if (row < m) g0 += gl[row] * v[row];
    return g0 + g1;
}

// G[row, :] . xs for the owned rows (0 for padding rows).
template <typename T, int NP, int MR, int LDG>
__device__ __forceinline__ void pdip_g_dot(const T *Gc, const T *xs, const bool (&rowvalid)[MR], int l,
                                           T (&out)[MR]) {
#pragma unroll
    for (int s = 0; s < MR; ++s) {
        T a0 = T(0), a1 = T(0);
        if (rowvalid[s]) {
            const T *gr = Gc + l + s * NP;
#pragma unroll
            for (int c = 0; c < NP; c += 2) {
                a0 += gr[c * LDG] * xs[c];
                a1 += gr[(c + 1) * LDG] * xs[c + 1];
            }
        }
        out[s] = a0 + a1;
    }
}

// Largest alpha in (0, 1] that keeps s + alpha ds and z + alpha dz positive,
// damped by frac (oracle/pdip_np.py:_step).
template <typename T, int NP, int MR>
__device__ __forceinline__ T pdip_step(const T (&sl)[MR], const T (&z)[MR], const T (&ds)[MR], const T (&dz)[MR],
                                       const bool (&rowvalid)[MR], T frac) {
    T a = Num<T>::inf();
#pragma unroll
    for (int s = 0; s < MR; ++s) {
        if (rowvalid[s] && ds[s] < T(0)) a = fmin(a, -sl[s] / ds[s]);
        if (rowvalid[s] && dz[s] < T(0)) a = fmin(a, -z[s] / dz[s]);
    }
    a = pdip_min<T, NP>(a);
    return fmin(T(1), frac * a);
}

// ---------------------------------------------------------------------------
// Phases B and C for one instance (all NP lanes of the group call this in
// lock step; the other groups of the warp run their own instances and the
// loop ends when every group of the warp is done).
//   in : Pc, Gc, hs in shared memory, qj = q_l; n, m real sizes (padding
//        variables have P = identity, q = 0, zero columns of G); pmin, a lower
//        bound of lambda_min(P) (w_u for a condensed MPC problem, mpc_qp.py:99-105)
//   out: x (component l of U), z (multipliers of the owned rows), status,
//        iterations
// Scratch: Lc [NP*LDL], xs [NP], dv [NP], wv [MP], tv [MP].
// ---------------------------------------------------------------------------
template <typename T, int NP, int MR, bool LS = false>
__device__ __forceinline__ void pdip_core(const T *Pc, T qj, const T *Gc, const T *hs, T *Lc, T *xs, T *dv, T *wv,
                                          T *tv, int m, int l, bool valid, int max_iter, T tol, bool polish, T pmin,
                                          T &x_out, T (&z_out)[MR], int &st_out, int &it_out) {
    using L = PdipLay<T, NP, MR>;
    constexpr int LDG = L::LDG, LDL = L::LDL;
    bool rowvalid[MR];
    T hrow[MR], sl[MR], z[MR], rp[MR], W[MR], ds[MR], dz[MR], rc[MR], gx[MR];
#pragma unroll
    for (int s = 0; s < MR; ++s) {
        const int row = l + s * NP;
        rowvalid[s] = row < m;
        hrow[s] = rowvalid[s] ? hs[row] : T(1);
        // start (pdip_np.py): u = 0, s = max(h, 1), z = 1 / s -- every complementarity product
        // starts at 1, so a row with a huge bound ("no bound" constants) does not inflate mu
        sl[s] = rowvalid[s] ? fmax(hrow[s], T(1)) : T(1);
        z[s] = rowvalid[s] ? T(1) / sl[s] : T(0);
        rp[s] = W[s] = ds[s] = dz[s] = rc[s] = gx[s] = T(0);
    }
    const T qscale = fmax(T(1), pdip_max<T, NP>(abs_(qj)));
    const T minv = m > 0 ? T(1) / (T)m : T(0);
    tol = fmax(tol, PdipNum<T>::tol_min);
    T x = T(0);
    int st = 1, it = 0;
    bool done = !valid;
    T Hrow[NP];

    while (true) {
        // residuals: r_d = P u + q + G'z, r_p = G u + s - h, mu = s'z / m
        xs[l] = x;
#pragma unroll
        for (int s = 0; s < MR; ++s)
            if (rowvalid[s]) tv[l + s * NP] = z[s];
        __syncwarp();
        T px0 = T(0), px1 = T(0);
#pragma unroll
        for (int c = 0; c < NP; c += 2) {
            px0 += Pc[c * LDL + l] * xs[c];
            px1 += Pc[(c + 1) * LDL + l] * xs[c + 1];
        }
        const T px = px0 + px1, gz = pdip_gt_dot<T, LDG>(Gc, tv, m, l);
        const T rd = px + qj + gz;
        pdip_g_dot<T, NP, MR, LDG>(Gc, xs, rowvalid, l, gx);
        // the primal residual is tested row by row, each against the size of its own terms
        // (one huge bound must not relax the test of the other rows): rpex <= 0 passes
        T comp = T(0), rpex = T(-1);
#pragma unroll
        for (int s = 0; s < MR; ++s) {
            rp[s] = rowvalid[s] ? gx[s] + sl[s] - hrow[s] : T(0);
            if (rowvalid[s]) {
                comp += sl[s] * z[s];
                const T scale = fmax(fmax(T(1), abs_(hrow[s])), fmax(abs_(gx[s]), sl[s]));
                rpex = fmax(rpex, abs_(rp[s]) - tol * scale);
            }
        }
        const T mu = pdip_sum<T, NP>(comp) * minv;
        const T rdmax = pdip_max<T, NP>(abs_(rd));
        const T rpmax = pdip_max<T, NP>(rpex);
        // relative criteria (pdip_np.py): each residual against the size of the terms it is
        // the sum of -- their rounding noise is the floor it can reach -- and the gap
        // against the objective.  All reductions first, then the (short-circuiting) test.
        const T dscale = fmax(qscale, pdip_max<T, NP>(fmax(abs_(px), abs_(gz))));
        const T obj = abs_(pdip_sum<T, NP>(x * (T(0.5) * px + qj)));
        const bool conv = rdmax <= tol * dscale && rpmax <= T(0) && mu <= tol * (T(1) + obj);
        if (!done && conv) {
            st = 0;
            done = true;
        }
        if (!done && it >= max_iter) done = true;  // st stays 1 (max_iter)
        if (__all_sync(FULL_MASK, done)) break;
        if (!done) ++it;
        __syncwarp();  // every lane has read z from tv

        // H = P + G'WG, W = z / s; predictor right-hand side -r_d - G'(W r_p - z)
#pragma unroll
        for (int s = 0; s < MR; ++s) {
            W[s] = rowvalid[s] ? z[s] / sl[s] : T(0);
            if (rowvalid[s]) {
                wv[l + s * NP] = W[s];
                tv[l + s * NP] = W[s] * rp[s] - z[s];
            }
        }
        __syncwarp();
        T rhs = -rd - pdip_build<T, NP, LDG, LDL>(Hrow, Pc, Gc, wv, tv, m, l);
        const bool spd = pdip_cholesky<T, NP, LDL>(Hrow, Lc, dv, l);
        __syncwarp();
        if (!done && !spd) {
            st = 3;  // numerical failure: H lost positive definiteness
            done = true;
        }
        if (LS) pdip_invert<T, NP, LDL>(Lc, dv, l);
        T du = LS ? pdip_solve_inv<T, NP, LDL>(Lc, xs, rhs, l) : pdip_solve<T, NP, LDL>(Lc, dv, rhs, l);
        xs[l] = du;
        __syncwarp();
        pdip_g_dot<T, NP, MR, LDG>(Gc, xs, rowvalid, l, gx);
#pragma unroll
        for (int s = 0; s < MR; ++s) {
            ds[s] = rowvalid[s] ? -rp[s] - gx[s] : T(0);
            dz[s] = rowvalid[s] ? -z[s] - W[s] * ds[s] : T(0);  // -(s z + z ds) / s
        }
        const T a_aff = pdip_step<T, NP, MR>(sl, z, ds, dz, rowvalid, T(1));
        T comp_aff = T(0);
#pragma unroll
        for (int s = 0; s < MR; ++s)
            if (rowvalid[s]) comp_aff += (sl[s] + a_aff * ds[s]) * (z[s] + a_aff * dz[s]);
        const T mu_aff = pdip_sum<T, NP>(comp_aff) * minv;
        const T ratio = mu > T(0) ? mu_aff / mu : T(0);
        const T sigmu = ratio * ratio * ratio * mu;

        // corrector: r_c = s z + ds_a dz_a - sigma mu
#pragma unroll
        for (int s = 0; s < MR; ++s) {
            rc[s] = rowvalid[s] ? sl[s] * z[s] + ds[s] * dz[s] - sigmu : T(0);
            if (rowvalid[s]) tv[l + s * NP] = W[s] * rp[s] - rc[s] / sl[s];
        }
        __syncwarp();
        rhs = -rd - pdip_gt_dot<T, LDG>(Gc, tv, m, l);
        du = LS ? pdip_solve_inv<T, NP, LDL>(Lc, xs, rhs, l) : pdip_solve<T, NP, LDL>(Lc, dv, rhs, l);
        xs[l] = du;
        __syncwarp();
        pdip_g_dot<T, NP, MR, LDG>(Gc, xs, rowvalid, l, gx);
#pragma unroll
        for (int s = 0; s < MR; ++s) {
            ds[s] = rowvalid[s] ? -rp[s] - gx[s] : T(0);
            dz[s] = rowvalid[s] ? -(rc[s] + z[s] * ds[s]) / sl[s] : T(0);
        }
        const T alpha = pdip_step<T, NP, MR>(sl, z, ds, dz, rowvalid, PdipNum<T>::step_frac);
        if (!done) {
            x += alpha * du;
#pragma unroll
            for (int s = 0; s < MR; ++s) {
                sl[s] += alpha * ds[s];
                z[s] += alpha * dz[s];
            }
        }
        __syncwarp();  // xs, tv are rewritten at the top of the loop
    }

    // ---- phase C: primal-dual active-set polish (pdip_np.py:_polish) --------
    // Proximal multiplier steps in residual form on the equality QP of the
    // guessed active set A; an accepted point is the exact solution on A.  A
    // rejected guess is corrected (negative multipliers leave, violated rows
    // enter) and the polish repeated.
    // Instances the iteration gave up on (cap, lost definiteness) are polished too: an
    // accepted point is a KKT-certified solution wherever the iterate came from.  They are
    // held to the absolute form of the acceptance test (`strict`: their iterate may be huge).
    if (__any_sync(FULL_MASK, polish && valid && m > 0)) {
        const T delta = PdipNum<T>::delta, dinv = T(1) / delta, eps = PdipNum<T>::polish_eps;
        const bool strict = st != 0;
        bool act[MR];
        T lam[MR], r2[MR];
        bool accepted = false;
        T up = x;
#pragma unroll
        for (int s = 0; s < MR; ++s) {
            act[s] = rowvalid[s] && z[s] > sl[s];
            lam[s] = act[s] ? z[s] : T(0);
            r2[s] = T(0);
        }
        for (int round = 0; round < PdipNum<T>::polish_rounds; ++round) {
            const bool need = polish && valid && m > 0 && !accepted;
            if (!__any_sync(FULL_MASK, need)) break;
            __syncwarp();
#pragma unroll
            for (int s = 0; s < MR; ++s) {
                lam[s] = act[s] ? lam[s] : T(0);
                if (rowvalid[s]) {
                    wv[l + s * NP] = act[s] ? dinv : T(0);
                    tv[l + s * NP] = T(0);
                }
            }
            __syncwarp();
            pdip_build<T, NP, LDG, LDL>(Hrow, Pc, Gc, wv, tv, m, l);
            const bool spd = pdip_cholesky<T, NP, LDL>(Hrow, Lc, dv, l);
            __syncwarp();
            if (LS) pdip_invert<T, NP, LDL>(Lc, dv, l);
            T rd = T(0), pxp = T(0), gtl = T(0);
            // Proximal steps move (up, lam) until the stationarity residual is as small as the
            // parity bar needs -- |r1| <= stat_eps lambda_min(P) |u| bounds the distance to the
            // exact point of this active set by stat_eps |u| -- or as small as rounding lets it
            // be, and the active rows hold; every pass starts by evaluating the residuals, the
            // last pass only evaluates them.
            T g_rd = T(0), g_ds = T(0), rd_tol = T(0), neg_tol = T(0);
            for (int step = 0;; ++step) {
                xs[l] = up;
#pragma unroll
                for (int s = 0; s < MR; ++s)
                    if (rowvalid[s]) tv[l + s * NP] = lam[s];
                __syncwarp();
                T px0 = T(0), px1 = T(0);
#pragma unroll
                for (int c = 0; c < NP; c += 2) {
                    px0 += Pc[c * LDL + l] * xs[c];
                    px1 += Pc[(c + 1) * LDL + l] * xs[c + 1];
                }
                pxp = px0 + px1;
                gtl = pdip_gt_dot<T, LDG>(Gc, tv, m, l);
                rd = pxp + qj + gtl;  // r1
                pdip_g_dot<T, NP, MR, LDG>(Gc, xs, rowvalid, l, gx);
                T act_res = T(-1);
#pragma unroll
                for (int s = 0; s < MR; ++s) {
                    gx[s] = rowvalid[s] ? gx[s] - hrow[s] : T(0);  // G up - h
                    if (act[s]) {
                        const T hs_ = fmax(T(1), abs_(hrow[s]));
                        act_res = fmax(act_res, abs_(gx[s]) - eps * (strict ? hs_ : fmax(hs_, abs_(gx[s] + hrow[s]))));
                    }
                }
                __syncwarp();
                g_rd = pdip_max<T, NP>(abs_(rd));
                g_ds = pdip_max<T, NP>(fmax(abs_(pxp), abs_(gtl)));
                const T uscale = fmax(T(1), pdip_max<T, NP>(abs_(up)));
                const T g_ar = pdip_max<T, NP>(act_res);
                rd_tol = fmax(PdipNum<T>::stat_eps * pmin * uscale,
                              T(64) * PdipNum<T>::macheps * (strict ? qscale : fmax(qscale, g_ds)));
                // a multiplier of -e on a row moves u by up to e |g| / lambda_min(P): held to the same bar
                neg_tol = PdipNum<T>::stat_eps * pmin * uscale;
                const bool tight = g_rd <= rd_tol && g_ar <= T(0);
                if (step == PdipNum<T>::polish_steps || __all_sync(FULL_MASK, tight || !need)) break;
#pragma unroll
                for (int s = 0; s < MR; ++s) {
                    r2[s] = act[s] ? gx[s] : T(0);
                    if (rowvalid[s]) tv[l + s * NP] = r2[s] * dinv;
                }
                __syncwarp();
                const T rhs = -(rd + pdip_gt_dot<T, LDG>(Gc, tv, m, l));
                const T du = LS ? pdip_solve_inv<T, NP, LDL>(Lc, xs, rhs, l) : pdip_solve<T, NP, LDL>(Lc, dv, rhs, l);
                __syncwarp();
                xs[l] = du;
                __syncwarp();
                pdip_g_dot<T, NP, MR, LDG>(Gc, xs, rowvalid, l, gx);
#pragma unroll
                for (int s = 0; s < MR; ++s) lam[s] += act[s] ? (r2[s] + gx[s]) * dinv : T(0);
                up += du;
                __syncwarp();
            }
            // accept a primal feasible point with non-negative multipliers, zero
            // residual on A and a small stationarity residual
            // (primal tests row by row against the row's own scale: *_ex <= 0 passes)
            T viol_ex = T(-1), act_ex = T(-1), worst_neg = T(0), zmax = T(0);
            bool finite = abs_(up) < Num<T>::inf();  // fmax / fmin drop NaNs: test for them explicitly
#pragma unroll
            for (int s = 0; s < MR; ++s) {
                if (rowvalid[s]) {
                    finite = finite && abs_(lam[s]) < Num<T>::inf() && abs_(gx[s]) < Num<T>::inf();
                    const T hs_ = fmax(T(1), abs_(hrow[s]));
                    const T ps = strict ? hs_ : fmax(hs_, abs_(gx[s] + hrow[s]));  // |G up|
                    viol_ex = fmax(viol_ex, gx[s] - eps * ps);
                    if (act[s]) act_ex = fmax(act_ex, abs_(gx[s]) - eps * ps);
                    worst_neg = fmax(worst_neg, -lam[s]);
                    zmax = fmax(zmax, abs_(lam[s]));
                }
            }
            // (every reduction is a shuffle sequence the whole warp must enter: evaluate
            // them all before combining -- inside a short-circuited `&&` the groups of a
            // warp would leave the chain at different links and the device deadlocks)
            const T zscale = fmax(T(1), pdip_max<T, NP>(zmax));
            const T g_viol = pdip_max<T, NP>(viol_ex), g_act = pdip_max<T, NP>(act_ex);
            const T g_neg = pdip_max<T, NP>(worst_neg);
            const T g_bad = pdip_max<T, NP>((finite && abs_(rd) < Num<T>::inf()) ? T(0) : T(1));
            const bool ok = spd && g_bad == T(0) && g_viol <= T(0) && g_act <= T(0) &&
                            g_neg <= fmax(neg_tol, T(64) * PdipNum<T>::macheps * zscale) && g_rd <= rd_tol;
            if (need && ok) {
                accepted = true;
                st = 0;
                x = up;
#pragma unroll
                for (int s = 0; s < MR; ++s) z[s] = lam[s];
            }
            // correct the guess by ONE row: the most violated row enters; if nothing is violated the
            // most negative multiplier leaves (changing many rows at once can cycle)
            T vbest = -Num<T>::inf(), lbest = Num<T>::inf();
            int vslot = 0, lslot = 0;
#pragma unroll
            for (int s = 0; s < MR; ++s) {
                if (rowvalid[s]) {
                    const T hs_ = fmax(T(1), abs_(hrow[s]));
                    const T ps = strict ? hs_ : fmax(hs_, abs_(gx[s] + hrow[s]));
                    const T v = gx[s] / ps;
                    if (!act[s] && v > vbest) {
                        vbest = v;
                        vslot = s;
                    }
                    if (act[s] && lam[s] < lbest) {
                        lbest = lam[s];
                        lslot = s;
                    }
                }
            }
            const T gv = pdip_max<T, NP>(vbest), gl = pdip_min<T, NP>(lbest);
            const unsigned lane_ = threadIdx.x & 31u;
            const unsigned segm = (NP == 32) ? FULL_MASK : (((1u << (NP & 31)) - 1u) << ((lane_ / NP) * NP));
            const unsigned bv = __ballot_sync(FULL_MASK, vbest == gv) & segm;
            const unsigned bl = __ballot_sync(FULL_MASK, lbest == gl) & segm;
            const bool enter = gv > eps && lane_ == (unsigned)(__ffs(bv) - 1);
            const bool leave = !(gv > eps) && gl < T(0) && lane_ == (unsigned)(__ffs(bl) - 1);
#pragma unroll
            for (int s = 0; s < MR; ++s) {
                if (enter && s == vslot) act[s] = true;
                if (leave && s == lslot) act[s] = false;
            }
        }
        __syncwarp();
    }
    x_out = x;
#pragma unroll
    for (int s = 0; s < MR; ++s) z_out[s] = z[s];
    st_out = st;
    it_out = it;
}

// ---------------------------------------------------------------------------
// The fused kernel: stage, condense, interior point, outputs.
// ---------------------------------------------------------------------------
template <typename T, int NP, int MR, bool LS>  // @phase pdip kernel
__global__ void __launch_bounds__(256, 1) mpc_pdip_kernel(const SolveParams p, int polish) {
    using L = PdipLay<T, NP, MR>;
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
    T *Gc = wk + p.g_off;
    T *hs = wk + L::oH;
    T *Lc = wk + L::oL;
    T *Pc = wk + L::oP;
    T *xs = wk + L::oV;
    T *dv = xs + NP;
    T *wv = wk + L::oW;
    T *tv = wv + L::MP;

    const T *in[OP_COUNT];
#pragma unroll
    for (int o = 0; o < OP_COUNT; ++o) {
        const OperandView &v = p.op[o];
        in[o] = v.ptr ? inbase + v.smem_off + (v.per_instance ? (valid ? iic : 0) * v.sz : 0) : nullptr;
    }

    // phase A: dense G (p.toeplitz = 0), h, row l of P, q_l
    T qj;
    {
        T Prow[NP];
        condense_dispatch<T, NP, MR, false>(p, in, Gc, static_cast<T *>(nullptr), hs, Lc, Gc, wk + p.scr_off, l,
                                            Prow, qj, inst, valid);
#pragma unroll
        for (int c = 0; c < NP; ++c) Pc[c * L::LDL + l] = Prow[c];  // P is symmetric
    }
    __syncwarp();

    T x, z[MR];
    int st, it;
    pdip_core<T, NP, MR, LS>(Pc, qj, Gc, hs, Lc, xs, dv, wv, tv, m, l, valid, p.max_iter, (T)p.tol, polish != 0,
                             (T)p.w_u, x, z, st, it);

    // phase D: outputs
    {
        const unsigned segmask = (NP == 32) ? FULL_MASK : (((1u << (NP & 31)) - 1u) << (sub * NP));
        const unsigned nonfinite = __ballot_sync(FULL_MASK, !(abs_(x) < Num<T>::inf())) & segmask;
        if (st == 0 && nonfinite) st = 3;
    }
    if (valid) {
        const T xo = (st == 0) ? x : Num<T>::nan();
        if (l < n && p.U) static_cast<T *>(p.U)[(size_t)inst * n + l] = xo;
        if (l == 0) {
            if (p.status) p.status[inst] = st;
            if (p.iters) p.iters[inst] = it;
        }
        if (p.Z) {
            T *Zb = static_cast<T *>(p.Z) + (size_t)inst * m;
#pragma unroll
            for (int s = 0; s < MR; ++s)
                if (l + s * NP < m) Zb[l + s * NP] = (st == 0) ? z[s] : T(0);
        }
    }
}

}  // namespace qpmpc
