// mpc_factor.cuh -- the shared-model fast path: factor ONCE per model, then per solve only the
// vectors change.
//
// This is the batched form of what a caller of the reference does when it keeps one MPCQP and,
// between control cycles, only calls update_cost_vector / update_constraint_vector
// (qpmpc/mpc_qp.py:129-163; usage: examples/wheeled_inverted_pendulum.py:99-118 rebuilds
// everything per cycle although only x0, the goal and the targets moved).  When A, B, C, D are
// shared by the batch, P, G, the Cholesky factor L, M = G L^-T and the maps
//     q = Fx x0 - Fg goal - Ft targets        (Fx = w_t psi_N' phi_N + w_x Psi' Phi,
//                                              Fg = w_t psi_N',  Ft = w_x Psi';  mpc_qp.py:139-149)
//     h = e - Hx x0                           (Hx = Cbar Phi;                    mpc_qp.py:161-163)
// are the same for every instance and every cycle.  mpc_factor_kernel computes them once from
// the condensed fields (the MPCQP of the model) into a RECORD in device memory; the solve kernel
// (mpc_solve_kernel<..., PRE = true>) then stages the record into shared memory once per CTA and
// starts every instance at "q, h, t = L^-1 q, violations": no recursion, no Cholesky, no
// substitutions per instance.  Paired rows (desc.paired), n <= 32.
#pragma once

#include "mpc_common.cuh"

namespace qpmpc {

// Offsets (elements of T) inside the record.  By columns / by lane so that a lane group reads
// consecutive addresses: X[t * NP + l].
struct FactorLay {
    int NP, LDL, LDG;
    int oL, oDv, oLinv, oM, oRowc, oG, oFx, oFg, oHx, oFt, oLinvT, oMT, LDM, total;
};
__host__ __device__ inline int up4(int v) { return (v + 3) / 4 * 4; }
__host__ __device__ inline FactorLay factor_layout(int NP, int nx, int N, bool has_ft) {
    FactorLay F;
    F.NP = NP;
    F.LDL = NP + 2;
    F.LDG = 2 * NP + 1;
    int o = 0;
    F.oL = o, o += up4(NP * F.LDL);      // L by columns: L[c * LDL + row]
    F.oDv = o, o += NP;                  // 1 / L_cc (NaN if P is not positive definite)
    F.oLinv = o, o += NP * NP;           // L^-1 by columns: Linv[k * NP + row]
    F.oM = o, o += NP * NP;              // stored rows of M = G+ L^-T by columns: M[c * NP + srow]
    F.oRowc = o, o += 3 * NP;            // |G_i|, 1 / |G_i|, |M_i|^2 per stored row
    F.oG = o, o += up4(NP * F.LDG);      // G (all m rows) by columns: G[c * LDG + row]
    F.oFx = o, o += nx * NP;             // Fx[t * NP + l]
    F.oFg = o, o += nx * NP;
    F.oHx = o, o += nx * NP;             // Hx of the stored (+) rows: Hx[t * NP + srow]
    F.oFt = o, o += has_ft ? N * nx * NP : 0;
    // for the recovery of x = -L^-T (t + M_A' lambda) by two short products instead of two
    // substitutions: L^-1 by ROWS (LinvT[k * NP + l] = (L^-1)_kl) and the stored rows of M by rows
    F.LDM = NP + 1;
    F.oLinvT = o, o += NP * NP;
    F.oMT = o, o += up4(NP * F.LDM);     // MT[srow * LDM + c] = M[srow][c]
    F.total = up4(o);
    return F;
}

struct FactorParams {
    int N, nx, nu, nc, n, m, NP;
    int has_ft, q_wt, q_wx;
    double w_t, w_x;
    // condensed fields of the model (row-major, as qpmpc_b200_condense writes them)
    const void *P, *G, *Phi, *Psi, *phi_last, *psi_last;
    const void *C;  // shared C (nullptr if absent), per-step stride stepC
    int stepC;
    void *record;
};

constexpr int FACTOR_SMEM_BYTES = 2 * 32 * 33 * 8 + 16;

// One CTA.  Everything is small (n <= 32): plain loops over global memory, barriers between steps.
template <typename T>  // @phase factor kernel
__global__ void __launch_bounds__(128) mpc_factor_kernel(const FactorParams p) {
    const int tid = threadIdx.x, nt = blockDim.x;
    const int n = p.n, m = p.m, nx = p.nx, N = p.N, NP = p.NP, half = p.nc / 2, mp = m / 2;
    const FactorLay F = factor_layout(NP, nx, N, p.has_ft != 0);
    T *R = static_cast<T *>(p.record);
    const T *P = static_cast<const T *>(p.P), *G = static_cast<const T *>(p.G);
    const T *Phi = static_cast<const T *>(p.Phi), *Psi = static_cast<const T *>(p.Psi);
    const T *phiN = static_cast<const T *>(p.phi_last), *psiN = static_cast<const T *>(p.psi_last);
    const T *C = static_cast<const T *>(p.C);
    extern __shared__ __align__(16) unsigned char smem_raw[];  // FACTOR_SMEM_BYTES
    T *Ls = reinterpret_cast<T *>(smem_raw);  // L (lower), row-major, ld 33
    T *Li = Ls + 32 * 33;                     // L^-1 (lower)
    int &bad = *reinterpret_cast<int *>(Li + 32 * 33);
    for (int i = tid; i < F.total; i += nt) R[i] = T(0);
    for (int i = tid; i < 32 * 33; i += nt) Ls[i] = Li[i] = T(0);
    if (tid == 0) bad = 0;
    __syncthreads();
    // padding variables: identity (as the solve kernels pad P)
    for (int i = tid; i < NP * NP; i += nt) {
        const int r = i / NP, c = i - r * NP;
        Ls[r * 33 + c] = (r < n && c < n) ? P[r * n + c] : (r == c ? T(1) : T(0));
    }
    __syncthreads();
    // Cholesky, in place (lower triangle)
    for (int c = 0; c < NP; ++c) {
        if (tid == 0) {
            const T piv = Ls[c * 33 + c];
            if (!(piv > T(0))) bad = 1;
            Ls[c * 33 + c] = sqrt_(piv);
        }
        __syncthreads();
        const T inv = T(1) / Ls[c * 33 + c];
        for (int i = c + 1 + tid; i < NP; i += nt) Ls[i * 33 + c] *= inv;
        __syncthreads();
        for (int idx = tid; idx < (NP - c - 1) * (NP - c - 1); idx += nt) {
            const int i = c + 1 + idx / (NP - c - 1), j = c + 1 + idx % (NP - c - 1);
            if (j <= i) Ls[i * 33 + j] -= Ls[i * 33 + c] * Ls[j * 33 + c];
        }
        __syncthreads();
    }
    // L^-1: thread c solves L y = e_c
    for (int c = tid; c < NP; c += nt) {
        for (int i = c; i < NP; ++i) {
            T s = (i == c) ? T(1) : T(0);
            for (int k = c; k < i; ++k) s -= Ls[i * 33 + k] * Li[k * 33 + c];
            Li[i * 33 + c] = s / Ls[i * 33 + i];
        }
    }
    __syncthreads();
    for (int i = tid; i < NP * NP; i += nt) {
        const int r = i / NP, c = i - r * NP;  // entry (r, c) of L and of L^-1
        if (c <= r) {
            R[F.oL + c * F.LDL + r] = Ls[r * 33 + c];
            R[F.oLinv + c * NP + r] = Li[r * 33 + c];
            R[F.oLinvT + r * NP + c] = Li[r * 33 + c];
        }
    }
    for (int c = tid; c < NP; c += nt) R[F.oDv + c] = bad ? Num<T>::nan() : T(1) / Ls[c * 33 + c];
    // G by columns (all rows); stored (+) rows of M = G L^-T with their norms
    for (int i = tid; i < m * n; i += nt) {
        const int r = i / n, c = i - r * n;
        R[F.oG + c * F.LDG + r] = G[r * n + c];
    }
    for (int srow = tid; srow < mp; srow += nt) {
        const int k = srow / half, r = srow - k * half;
        const T *g = G + (size_t)(k * p.nc + r) * n;
        T g2 = T(0), m2 = T(0);
        for (int c = 0; c < n; ++c) g2 += g[c] * g[c];
        for (int c = 0; c < NP; ++c) {
            T s = T(0);  // M[srow][c] = sum_k G[srow][k] Linv[c][k]
            for (int kk = 0; kk <= c && kk < n; ++kk) s += g[kk] * Li[c * 33 + kk];
            R[F.oM + c * NP + srow] = s;
            R[F.oMT + srow * F.LDM + c] = s;
            m2 += s * s;
        }
        R[F.oRowc + srow] = sqrt_(g2);
        R[F.oRowc + NP + srow] = g2 > T(0) ? T(1) / sqrt_(g2) : T(1e30);
        R[F.oRowc + 2 * NP + srow] = m2;
    }
    // Fx, Fg, Ft
    const T w_t = (T)p.w_t, w_x = (T)p.w_x;
    for (int i = tid; i < nx * n; i += nt) {
        const int t = i / n, l = i - t * n;
        T fx = T(0), fg = T(0);
        if (p.q_wt) {
            for (int s = 0; s < nx; ++s) fx += w_t * psiN[s * n + l] * phiN[s * nx + t];
            fg = w_t * psiN[t * n + l];
        }
        if (p.q_wx)
            for (int j = 0; j < N * nx; ++j) fx += w_x * Psi[(size_t)j * n + l] * Phi[j * nx + t];
        R[F.oFx + t * NP + l] = fx;
        R[F.oFg + t * NP + l] = fg;
    }
    if (p.has_ft)
        for (int i = tid; i < N * nx * n; i += nt) {
            const int j = i / n, l = i - j * n;
            R[F.oFt + j * NP + l] = w_x * Psi[(size_t)j * n + l];
        }
    // Hx of the stored rows: C_k[r, :] phi_k
    if (C)
        for (int i = tid; i < mp * nx; i += nt) {
            const int srow = i / nx, t = i - srow * nx;
            const int k = srow / half, r = srow - k * half;
            const T *Ck = C + (size_t)k * p.stepC;
            T s = T(0);
            for (int u = 0; u < nx; ++u) s += Ck[r * nx + u] * Phi[(size_t)(k * nx + u) * nx + t];
            R[F.oHx + t * NP + srow] = s;
        }
}

}  // namespace qpmpc
