// mpc_cta_kernel.cuh -- one CTA per MPC instance, for problems beyond the
// warp kernel (n = N*nu > 32 or m = N*nc > 4 n; the N = 64 horizon-sweep
// point has n = 64, m = 128).  Same algorithm as mpc_kernels.cuh (condense,
// Cholesky, Goldfarb-Idnani dual active set with Householder adds and an
// explicit R^-1) but every matrix lives in shared memory, all loops have
// run-time bounds and the CTA's threads split rows between them:
//
//   M   m x ld   G, then M = G J in place (row-major, row r owned by one thread)
//   J   n x ld   J = L^-T                  (row-major, row l owned by one thread)
//   PL  n x ld   P, then L (in place), then R by columns
//   Ri  n x ld   condensing scratch, then R^-1 by columns
//   vectors      h, viol, vtol, ginv, |M_i|^2, G z  [m];  q, x, z, d, d2, lam, r,
//                1/L_kk, t, cand, cs, sn [n];  aidx [n], active flags [m]
//
// ld = n | 1 (odd) so that threads walking their own rows hit distinct banks.
// Reference citations as in mpc_kernels.cuh (qpmpc/mpc_qp.py:53-105,139-149;
// qpmpc/solve_mpc.py:43).
#pragma once

#include "mpc_common.cuh"

namespace qpmpc {

struct CtaLay {
    int ld, oM, oJ, oPL, oRi, oV, total;  // offsets in elements of T
    int o_hs, o_viol, o_vtol, o_ginv, o_mn2, o_gz;
    int o_q, o_x, o_z, o_dd, o_d2, o_lam, o_rv, o_dv, o_tq, o_cand, o_cs, o_sn;
    int o_aidx, o_actf, o_red, o_sc;  // int / int / 16 x 8-byte reduction slots / 8 scalars
};

__host__ __device__ inline int even_up(int v) { return (v + 1) & ~1; }

// `esz` = sizeof(T).  Everything is expressed in elements of T; the int arrays
// are rounded up to whole elements.
__host__ __device__ inline CtaLay cta_layout(int n, int m, int nx, int esz) {
    CtaLay L;
    L.ld = n | 1;
    const int mat = even_up(n * L.ld);
    // generic condensing: psi[2][nx][n], xbar[2][nx], phi[2][nx][nx]; LTI path: W[nx][n], XB[N+1][nx] (N <= n)
    const int scratch = even_up(2 * nx * n + 2 * nx + 3 * nx * nx + nx);
    L.oM = 0;
    L.oJ = L.oM + even_up(m * L.ld);
    L.oPL = L.oJ + mat;
    L.oRi = L.oPL + mat;
    L.oV = L.oRi + (mat > scratch ? mat : scratch);
    int o = L.oV;
    const int me = even_up(m > 0 ? m : 1), ne = even_up(n);
    L.o_hs = o, o += me;
    L.o_viol = o, o += me;
    L.o_vtol = o, o += me;
    L.o_ginv = o, o += me;
    L.o_mn2 = o, o += me;
    L.o_gz = o, o += me;
    L.o_q = o, o += ne;
    L.o_x = o, o += ne;
    L.o_z = o, o += ne;
    L.o_dd = o, o += ne;
    L.o_d2 = o, o += ne;
    L.o_lam = o, o += ne;
    L.o_rv = o, o += ne;
    L.o_dv = o, o += ne;
    L.o_tq = o, o += ne;
    L.o_cand = o, o += ne;
    L.o_cs = o, o += ne;
    L.o_sn = o, o += ne;
    L.o_aidx = o, o += even_up((ne * 4 + esz - 1) / esz);
    L.o_actf = o, o += even_up((me * 4 + esz - 1) / esz);
    L.o_red = o, o += even_up((16 * 8 + esz - 1) / esz);
    L.o_sc = o, o += 8;
    L.total = o;
    return L;
}

// ---- block-wide reductions over 64-bit keys ---------------------------------
__device__ __forceinline__ unsigned long long block_reduce_max(unsigned long long key, unsigned long long *red,
                                                               int parity) {
    // `red` has two sets of 8 slots used alternately, so one barrier per call is
    // enough (a warp can only be one call ahead of the slowest reader).
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const unsigned long long o = __shfl_xor_sync(FULL_MASK, key, off);
        key = o > key ? o : key;
    }
    unsigned long long *slot = red + (parity & 1) * 8;
    const int warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    if ((threadIdx.x & 31) == 0) slot[warp] = key;
    __syncthreads();
    unsigned long long r = 0ull;
    for (int w = 0; w < nw; ++w) r = slot[w] > r ? slot[w] : r;
    return r;
}

// Row index in the low 12 bits (m <= 4096); larger score wins, ties -> lower row.
__device__ __forceinline__ unsigned long long wide_key(double s, int idx) {
    return ((unsigned long long)__double_as_longlong(s) & ~4095ull) | (unsigned)(4095 - idx);
}
__device__ __forceinline__ unsigned long long wide_key(float s, int idx) {
    return ((unsigned long long)__float_as_uint(s) << 32) | (unsigned)(4095 - idx);
}

template <typename T>
__device__ __forceinline__ const T *operand(const SolveParams &p, int o, long long inst) {
    const OperandView &v = p.op[o];
    if (!v.ptr) return nullptr;
    return static_cast<const T *>(v.ptr) + (v.per_instance ? (size_t)inst * v.sz : 0);
}

// ---------------------------------------------------------------------------
// Condensing by a whole CTA (qpmpc/mpc_qp.py:53-105, 139-149).  Thread c < n
// rolls column c of psi_k through the scratch region; G goes to M (row-major),
// P to PL (full symmetric matrix), q and h to their vectors.
// ---------------------------------------------------------------------------
// ---------------------------------------------------------------------------
// Condensing of a time-invariant model (A, B, C constant over the horizon; D, e
// may vary) without the N-step barrier loop.  psi_k[:, c] = A^(k-1-j) B[:, jj]
// (c = j nu + jj) depends on e = k nu - 1 - c only, so two serial chains, run
// by one thread each, produce everything the rest needs:
//   W[:, e] = A^d B[:, jj],  e = d nu + (nu - 1 - jj)      (thread 0)
//   XB[k]  = A^k x0,  k = 0..N                             (thread 32)
// and then, in parallel over entries,
//   G[(k, r), c] = C_r W[:, k nu - 1 - c]  (c < k nu),  D_k on block column k,
//   h[(k, r)]    = e_k[r] - C_r XB[k],
//   P = w_u I + w_t psi_N' psi_N + w_x sum_k psi_k' psi_k  (the last term from
//       prefix sums along the Toeplitz diagonals, nu = 1), q likewise.
// Scratch: W, XB in the R^-1 region, the prefix-sum table in the J region.
// ---------------------------------------------------------------------------
template <typename T>  // @phase CTA condense (time-invariant model)
__device__ void cta_condense_lti(const SolveParams &p, const CtaLay &L, T *mat, T *vec, long long inst) {
    const int tid = threadIdx.x, nt = blockDim.x;
    const int nx = p.nx, nu = p.nu, nc = p.nc, N = p.N, n = p.n, m = p.m, ld = L.ld;
    T *M = mat + L.oM, *PL = mat + L.oPL, *Tt = mat + L.oJ;
    T *W = mat + L.oRi;           // [nx][n]
    T *XB = W + nx * n;           // [N + 1][nx]
    T *hs = vec + L.o_hs, *q = vec + L.o_q;
    const T *A = operand<T>(p, OP_A, inst), *B = operand<T>(p, OP_B, inst);
    const T *C = operand<T>(p, OP_C, inst), *D = operand<T>(p, OP_D, inst);
    const T *e = operand<T>(p, OP_E, inst), *x0 = operand<T>(p, OP_X0, inst);
    const T *goal = operand<T>(p, OP_GOAL, inst), *tgt = operand<T>(p, OP_TGT, inst);
    const T w_t = (T)p.w_t, w_x = (T)p.w_x;
    // Thread jj < nu runs the chain of input column jj, thread 32 the free
    // response; vectors stay in registers (nx <= 8), A is read from shared memory.
    T *As = XB + (N + 1) * nx;
    for (int i = tid; i < nx * nx; i += nt) As[i] = A[i];
    __syncthreads();
    const int chain = tid < nu ? tid : ((tid == 32 || (nt <= 32 && tid == nu)) ? nu : -1);
    if (chain >= 0) {
        T v[8], w[8];
#pragma unroll
        for (int t = 0; t < 8; ++t) v[t] = t < nx ? (chain < nu ? B[t * nu + chain] : x0[t]) : T(0);
        const int steps = chain < nu ? N - 1 : N;
        for (int d = 0; d <= steps; ++d) {
            // store the current vector, then advance it by A
            if (chain < nu) {
#pragma unroll
                for (int t = 0; t < 8; ++t)
                    if (t < nx) W[t * n + d * nu + (nu - 1 - chain)] = v[t];
            } else {
#pragma unroll
                for (int t = 0; t < 8; ++t)
                    if (t < nx) XB[d * nx + t] = v[t];
            }
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                T acc = T(0);
#pragma unroll
                for (int s2 = 0; s2 < 8; ++s2)
                    if (t < nx && s2 < nx) acc += As[t * nx + s2] * v[s2];
                w[t] = acc;
            }
#pragma unroll
            for (int t = 0; t < 8; ++t) v[t] = w[t];
        }
    }
    __syncthreads();
    // G and h
    for (int idx = tid; idx < m * n; idx += nt) {
        const int row = idx / n, c = idx - row * n;
        const int k = row / nc, r = row - k * nc, kb = k * nu;
        T g = T(0);
        if (c < kb) {
            if (C)
                for (int t = 0; t < nx; ++t) g += C[r * nx + t] * W[t * n + kb - 1 - c];
        } else if (D && c - kb < nu) {
            g = D[k * p.op[OP_D].step + r * nu + c - kb];
        }
        M[row * ld + c] = g;
    }
    for (int row = tid; row < m; row += nt) {
        const int k = row / nc, r = row - k * nc;
        T hv = e[k * p.op[OP_E].step + r];
        if (C)
            for (int t = 0; t < nx; ++t) hv -= C[r * nx + t] * XB[k * nx + t];
        hs[row] = hv;
    }
    // prefix sums T_d[u] = sum_{a <= u} v_a . v_(a+d) for the stage-cost Hessian (nu = 1)
    if (p.has_wx)
        for (int d = tid; d < N; d += nt) {
            T acc = T(0);
            for (int u = 0; u + d < N; ++u) {
                for (int t = 0; t < nx; ++t) acc += W[t * n + u] * W[t * n + u + d];
                Tt[d * ld + u] = acc;
            }
        }
    __syncthreads();
    // P (full symmetric matrix) and q
    for (int idx = tid; idx < n * n; idx += nt) {
        const int i = idx / n, j = idx - i * n;
        T acc = (i == j) ? (T)p.w_u : T(0);
        if (p.has_wt) {
            T a = T(0);
            for (int t = 0; t < nx; ++t) a += (w_t * W[t * n + n - 1 - i]) * W[t * n + n - 1 - j];
            acc += a;
        }
        if (p.has_wx) {
            const int hi = i > j ? i : j, d = i > j ? i - j : j - i;
            if (hi <= N - 2) acc += w_x * Tt[d * ld + (N - 2 - hi)];
        }
        PL[i * ld + j] = acc;
    }
    for (int c = tid; c < n; c += nt) {
        T acc = T(0);
        if (p.q_wt)
            for (int t = 0; t < nx; ++t) acc += (w_t * W[t * n + n - 1 - c]) * (XB[N * nx + t] - goal[t]);
        if (p.q_wx)
            for (int k = c / nu + 1; k < N; ++k) {
                const int e1 = k * nu - 1 - c;
                for (int t = 0; t < nx; ++t) acc += (w_x * W[t * n + e1]) * (XB[k * nx + t] - tgt[k * nx + t]);
            }
        q[c] = acc;
    }
    __syncthreads();
}

template <typename T, bool DUMP>  // @phase CTA condense
__device__ void cta_condense(const SolveParams &p, const CtaLay &L, T *mat, T *vec, long long inst) {
    const int tid = threadIdx.x, nt = blockDim.x;
    const int nx = p.nx, nu = p.nu, nc = p.nc, N = p.N, n = p.n, ld = L.ld;
    T *M = mat + L.oM, *PL = mat + L.oPL, *scr = mat + L.oRi;
    T *hs = vec + L.o_hs, *q = vec + L.o_q;
    T *psi = scr;                   // [2][nx][n]
    T *xbar = psi + 2 * nx * n;     // [2][nx]
    T *phi = xbar + 2 * nx;         // [2][nx][nx]
    const T *A = operand<T>(p, OP_A, inst), *B = operand<T>(p, OP_B, inst);
    const T *C = operand<T>(p, OP_C, inst), *D = operand<T>(p, OP_D, inst);
    const T *e = operand<T>(p, OP_E, inst), *x0 = operand<T>(p, OP_X0, inst);
    const T *goal = operand<T>(p, OP_GOAL, inst), *tgt = operand<T>(p, OP_TGT, inst);
    const T w_t = (T)p.w_t, w_x = (T)p.w_x;

    for (int i = tid; i < n * ld; i += nt) PL[i] = T(0);
    for (int i = tid; i < nx * n; i += nt) psi[i] = T(0);
    for (int i = tid; i < n; i += nt) q[i] = T(0);
    for (int i = tid; i < nx; i += nt) xbar[i] = x0[i];
    if (DUMP)
        for (int i = tid; i < nx * nx; i += nt) phi[i] = (i / nx == i % nx) ? T(1) : T(0);
    __syncthreads();
    for (int i = tid; i < n; i += nt) PL[i * ld + i] = (T)p.w_u;
    int cur = 0;
    for (int k = 0; k <= N; ++k) {
        const T *ps = psi + cur * nx * n;
        const T *xb = xbar + cur * nx;
        const bool last = (k == N);
        const T w = last ? w_t : w_x;
        const bool inP = !p.skip_P && (last ? p.has_wt : p.has_wx);
        const bool inq = last ? p.q_wt : p.q_wx;
        const T *ref = last ? goal : (tgt ? tgt + k * nx : nullptr);
        if (!last) {
            const T *Ck = C ? C + k * p.op[OP_C].step : nullptr;
            const T *Dk = D ? D + k * p.op[OP_D].step : nullptr;
            const T *ek = e ? e + k * p.op[OP_E].step : nullptr;
            for (int idx = tid; idx < nc * n; idx += nt) {
                const int r = idx / n, c = idx - r * n;
                T g = T(0);
                if (Ck)
                    for (int t = 0; t < nx; ++t) g += Ck[r * nx + t] * ps[t * n + c];
                const int jj = c - k * nu;
                if (Dk && jj >= 0 && jj < nu) g += Dk[r * nu + jj];
                M[(k * nc + r) * ld + c] = g;
            }
            for (int r = tid; r < nc; r += nt) {
                T hv = ek[r];
                if (Ck)
                    for (int t = 0; t < nx; ++t) hv -= Ck[r * nx + t] * xb[t];
                hs[k * nc + r] = hv;
            }
            if (DUMP) {
                if (p.Psi)
                    for (int idx = tid; idx < nx * n; idx += nt)
                        static_cast<T *>(p.Psi)[((size_t)inst * N * nx + (size_t)k * nx) * n + idx] = ps[idx];
                if (p.Phi)
                    for (int idx = tid; idx < nx * nx; idx += nt)
                        static_cast<T *>(p.Phi)[((size_t)inst * N * nx + (size_t)k * nx) * nx + idx] =
                            phi[cur * nx * nx + idx];
            }
        }
        // cost terms: P += w psi' psi, q += w psi'(xb - ref)
        if (inP) {
            for (int i = tid >> 4; i < n; i += nt >> 4) {
                for (int j = tid & 15; j < n; j += 16) {
                    T acc = T(0);
                    for (int t = 0; t < nx; ++t) acc += (w * ps[t * n + i]) * ps[t * n + j];
                    PL[i * ld + j] += acc;
                }
            }
        }
        if (inq) {
            for (int c = tid; c < n; c += nt) {
                T acc = T(0);
                for (int t = 0; t < nx; ++t) acc += (w * ps[t * n + c]) * (xb[t] - ref[t]);
                q[c] += acc;
            }
        }
        if (last) break;
        // advance: psi_{k+1} = A_k psi_k, block column k := B_k; xb_{k+1} = A_k xb_k
        const T *Ak = A + k * p.op[OP_A].step;
        const T *Bk = B + k * p.op[OP_B].step;
        T *pn = psi + (cur ^ 1) * nx * n;
        T *xn = xbar + (cur ^ 1) * nx;
        for (int idx = tid; idx < nx * n; idx += nt) {
            const int t = idx / n, c = idx - t * n;
            const int jj = c - k * nu;
            T acc = T(0);
            if (jj >= 0 && jj < nu) {
                acc = Bk[t * nu + jj];
            } else {
                for (int s = 0; s < nx; ++s) acc += Ak[t * nx + s] * ps[s * n + c];
            }
            pn[idx] = acc;
        }
        for (int t = tid; t < nx; t += nt) {
            T acc = T(0);
            for (int s = 0; s < nx; ++s) acc += Ak[t * nx + s] * xb[s];
            xn[t] = acc;
        }
        if (DUMP) {
            const T *ph = phi + cur * nx * nx;
            T *pnx = phi + (cur ^ 1) * nx * nx;
            for (int idx = tid; idx < nx * nx; idx += nt) {
                const int r = idx / nx, c = idx - r * nx;
                T acc = T(0);
                for (int s = 0; s < nx; ++s) acc += Ak[r * nx + s] * ph[s * nx + c];
                pnx[idx] = acc;
            }
        }
        __syncthreads();
        cur ^= 1;
    }
    if (DUMP) {
        const T *ps = psi + cur * nx * n;
        if (p.psi_last)
            for (int idx = tid; idx < nx * n; idx += nt)
                static_cast<T *>(p.psi_last)[(size_t)inst * nx * n + idx] = ps[idx];
        if (p.phi_last)
            for (int idx = tid; idx < nx * nx; idx += nt)
                static_cast<T *>(p.phi_last)[(size_t)inst * nx * nx + idx] = phi[cur * nx * nx + idx];
    }
    __syncthreads();
}

// Forward substitution L y = rhs on a row the calling thread owns (in place).
template <typename T>
__device__ __forceinline__ void row_fsolve(const T *PL, const T *dv, int ld, int n, T *row, int first) {
    for (int c = first; c < n; ++c) {
        const T *Lr = PL + c * ld;
        T a0 = row[c], a1 = T(0);
        int k = first;
        for (; k + 1 < c; k += 2) {
            a0 -= Lr[k] * row[k];
            a1 -= Lr[k + 1] * row[k + 1];
        }
        if (k < c) a0 -= Lr[k] * row[k];
        row[c] = (a0 + a1) * dv[c];
    }
}

// Where the matrices (M, J, L / R, R^-1) and the vectors of one CTA live.  Shapes whose matrices
// fit keep everything in shared memory; larger ones (n > 72 in fp64 with m = 2 n) keep the
// vectors there and the matrices in a slice of a global-memory workspace the launcher
// allocates (p.workspace, one slice of L.oV elements per CTA) -- slower, but no horizon is
// refused for size.  The vector offsets of CtaLay count from the start of the matrices.
template <typename T>
__device__ __forceinline__ void cta_bases(const SolveParams &p, const CtaLay &L, T *sm, T *&mat, T *&vec) {
    if (p.workspace) {
        mat = static_cast<T *>(p.workspace) + (size_t)blockIdx.x * L.oV;
        vec = sm - L.oV;
    } else {
        mat = vec = sm;
    }
}

template <typename T>
__global__ void __launch_bounds__(256) mpc_solve_cta_kernel(const SolveParams p) {  // @phase CTA prologue
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *sm = reinterpret_cast<T *>(smem_raw);
    const int tid = threadIdx.x, nt = blockDim.x;
    const int n = p.n, m = p.m;
    const CtaLay L = cta_layout(n, m, p.nx, (int)sizeof(T));
    const int ld = L.ld;
    T *mat, *vec;
    cta_bases<T>(p, L, sm, mat, vec);
    T *M = mat + L.oM, *J = mat + L.oJ, *PL = mat + L.oPL, *Rc = mat + L.oPL, *Ri = mat + L.oRi;
    T *hs = vec + L.o_hs, *viol = vec + L.o_viol, *vtol = vec + L.o_vtol, *ginv = vec + L.o_ginv;
    T *mn2 = vec + L.o_mn2, *gzv = vec + L.o_gz;
    T *q = vec + L.o_q, *x = vec + L.o_x, *z = vec + L.o_z, *dd = vec + L.o_dd, *d2 = vec + L.o_d2;
    T *lam = vec + L.o_lam, *rv = vec + L.o_rv, *dv = vec + L.o_dv, *tq = vec + L.o_tq;
    T *cand = vec + L.o_cand, *cs = vec + L.o_cs, *sn = vec + L.o_sn;
    int *aidx = reinterpret_cast<int *>(vec + L.o_aidx);
    int *actf = reinterpret_cast<int *>(vec + L.o_actf);
    unsigned long long *red = reinterpret_cast<unsigned long long *>(vec + L.o_red);
    T *sc = vec + L.o_sc;

    // one instance per CTA, or (workspace mode: the grid is bounded) a stride over the batch
    for (long long inst = blockIdx.x; inst < p.batch; inst += gridDim.x) {
    {
        const bool lti = p.op[OP_A].step == 0 && p.op[OP_B].step == 0 && (!p.op[OP_C].ptr || p.op[OP_C].step == 0);
        if (lti && p.nc > 0 && p.nx <= 8 && p.nu < 32 && (!p.has_wx || p.nu == 1) && p.toeplitz)
            cta_condense_lti<T>(p, L, mat, vec, inst);
        else
            cta_condense<T, false>(p, L, mat, vec, inst);
    }

    // ---- Cholesky in place on PL (lower triangle)  // @phase CTA cholesky
    bool spd = true;
    {
        const int ti = tid >> 4, tj = tid & 15, rows = nt >> 4;
        for (int c = 0; c < n; ++c) {
            // the diagonal entry keeps the pivot (nothing reads L_cc: 1/L_cc is in dv)
            const T piv = PL[c * ld + c];
            spd = spd && (piv > T(0));
            const T inv = rsqrt_(piv);
            for (int i = c + 1 + tid; i < n; i += nt) PL[i * ld + c] *= inv;
            if (tid == 0) dv[c] = inv;
            __syncthreads();
            for (int i = c + 1 + ti; i < n; i += rows) {
                const T lic = PL[i * ld + c];
                for (int j = c + 1 + tj; j <= i; j += 16) PL[i * ld + j] -= lic * PL[j * ld + c];
            }
            __syncthreads();
        }
    }
    // ---- J = L^-T rows, t = L^-1 q  // @phase CTA J, x, M
    for (int task = tid; task <= n; task += nt) {
        if (task < n) {
            T *row = J + task * ld;
            for (int c = 0; c < n; ++c) row[c] = (c == task) ? T(1) : T(0);
            row_fsolve<T>(PL, dv, ld, n, row, task);
        } else {
            for (int c = 0; c < n; ++c) tq[c] = q[c];
            row_fsolve<T>(PL, dv, ld, n, tq, 0);
        }
    }
    __syncthreads();
    for (int l = tid; l < n; l += nt) {
        const T *row = J + l * ld;
        T a = T(0);
        for (int c = l; c < n; ++c) a -= row[c] * tq[c];
        x[l] = a;
    }
    for (int r = tid; r < m; r += nt) actf[r] = 0;
    __syncthreads();
    // ---- violations, then M = G J in place (row solves)
    for (int r = tid; r < m; r += nt) {
        T *row = M + r * ld;
        T vi = T(0), g2 = T(0);
        for (int c = 0; c < n; ++c) {
            const T g = row[c];
            vi += g * x[c];
            g2 += g * g;
        }
        const T hi = hs[r];
        viol[r] = vi - hi;
        vtol[r] = Num<T>::viol_eps * (fmax(T(1), abs_(hi)) + sqrt_(g2));
        ginv[r] = g2 > T(0) ? rsqrt_(g2) : T(1e30);
        row_fsolve<T>(PL, dv, ld, n, row, 0);
        T m2 = T(0);
        for (int c = 0; c < n; ++c) m2 += row[c] * row[c];
        mn2[r] = m2;
    }
    __syncthreads();  // PL is dead: R takes its place

    // ---- dual active-set iteration  // @phase CTA active-set loop
    const T INF = Num<T>::inf();
    int na = 0, it = 0, st = spd ? 0 : 3, pidx = 0;
    bool cont = false;
    T lamp = T(0);
    while (st == 0 && m > 0) {
        if (!cont) {
            unsigned long long key = 0ull;
            for (int r = tid; r < m; r += nt) {
                if (!actf[r] && viol[r] > vtol[r]) {
                    const unsigned long long ks = wide_key(viol[r] * ginv[r], r);
                    key = ks > key ? ks : key;
                }
            }
            key = block_reduce_max(key, red, it);
            if (key == 0ull) break;  // primal feasible: optimal
            pidx = 4095 - (int)(key & 4095ull);
            lamp = T(0);
        }
        if (++it > p.max_iter) {
            st = 1;
            break;
        }
        // d = -(row p of M)
        for (int c = tid; c < n; c += nt) {
            const T dl = -M[pidx * ld + c];
            dd[c] = dl;
            d2[c] = (c >= na) ? dl : T(0);
            if (c == na) sc[2] = dl;
        }
        if (tid == 0) {
            sc[0] = viol[pidx];
            sc[1] = mn2[pidx];
        }
        __syncthreads();
        // z = J2 d2, G z = M2 d2, r = R^-1 d1, |d2|^2
        for (int task = tid; task < n + m; task += nt) {
            const T *row = task < n ? J + task * ld : M + (task - n) * ld;
            T a0 = T(0), a1 = T(0);
            int c = na;
            for (; c + 1 < n; c += 2) {
                a0 += row[c] * d2[c];
                a1 += row[c + 1] * d2[c + 1];
            }
            if (c < n) a0 += row[c] * d2[c];
            if (task < n)
                z[task] = a0 + a1;
            else
                gzv[task - n] = a0 + a1;
        }
        for (int l = tid; l < na; l += nt) {
            T a = T(0);
            for (int k = l; k < na; ++k) a += Ri[k * ld + l] * dd[k];
            rv[l] = a;
            cand[l] = a > T(0) ? lam[l] / a : INF;
        }
        __syncthreads();
        // |d2|^2 and the step-length minimum, per warp (every warp gets the same values)
        const int lane = tid & 31;
        T a2 = T(0);
        for (int c = na + lane; c < n; c += 32) a2 += d2[c] * d2[c];
        T t1 = INF;
        int lidx = 0;
        for (int l = lane; l < na; l += 32) {
            const T c = cand[l];
            if (c < t1) {
                t1 = c;
                lidx = l;
            }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            a2 += __shfl_xor_sync(FULL_MASK, a2, off);
            const T ot = __shfl_xor_sync(FULL_MASK, t1, off);
            const int ol = __shfl_xor_sync(FULL_MASK, lidx, off);
            if (ot < t1 || (ot == t1 && ol < lidx)) {
                t1 = ot;
                lidx = ol;
            }
        }
        // scalars published with d: nothing read below is written again before the closing barrier
        const T violp = sc[0], dn2 = sc[1];
        const bool zzero = !(a2 > Num<T>::dep_eps * dn2);
        const T t2 = zzero ? INF : violp / a2;
        if (t1 == INF && t2 == INF) {
            st = 2;  // infeasible
            break;
        }
        const T t = fmin(t1, t2);
        const bool full = !zzero && t2 <= t1;
        const T dna = na < n ? sc[2] : T(0);
        if (!zzero) {
            for (int l = tid; l < n; l += nt) x[l] += t * z[l];
            for (int r = tid; r < m; r += nt) viol[r] += t * gzv[r];
        }
        for (int l = tid; l < na; l += nt) lam[l] -= t * rv[l];
        lamp += t;
        if (full) {
            // Householder: H = I - tau v v', v = d2 - beta e_na, on columns >= na of J and M
            const T ainv = rsqrt_(a2);
            const T alpha = a2 * ainv;
            const T beta = (dna > T(0)) ? -alpha : alpha;
            const T binv = (dna > T(0)) ? -ainv : ainv;
            const T tau = T(1) / (a2 - beta * dna);
            const T vna = dna - beta;  // v differs from d2 in entry na only
            for (int task = tid; task < n + m; task += nt) {
                T *row = task < n ? J + task * ld : M + (task - n) * ld;
                T a0 = row[na] * vna, a1 = T(0);
                int c = na + 1;
                for (; c + 1 < n; c += 2) {
                    a0 += row[c] * d2[c];
                    a1 += row[c + 1] * d2[c + 1];
                }
                if (c < n) a0 += row[c] * d2[c];
                const T s = (a0 + a1) * tau;
                row[na] -= s * vna;
                for (c = na + 1; c < n; ++c) row[c] -= s * d2[c];
            }
            for (int l = tid; l <= na; l += nt) {
                if (l < na) {
                    Rc[na * ld + l] = dd[l];
                    Ri[na * ld + l] = -rv[l] * binv;
                } else {
                    Rc[na * ld + na] = beta;
                    Ri[na * ld + na] = binv;
                    lam[na] = lamp;
                    aidx[na] = pidx;
                    actf[pidx] = 1;
                }
            }
            ++na;
            cont = false;
            __syncthreads();
        } else {
            // drop the constraint at active position lidx; p stays the candidate
            const int nan_ = na - 1;
            __syncthreads();  // lam is up to date
            if (tid == 0) {
                actf[aidx[lidx]] = 0;
                for (int l = lidx; l < nan_; ++l) {
                    lam[l] = lam[l + 1];
                    aidx[l] = aidx[l + 1];
                }
            }
            // shift columns lidx+1.. of R one to the left (column l+1 -> l), one column per thread
            for (int l = lidx; l < nan_; ++l) {
                __syncthreads();
                for (int row = tid; row <= l + 1; row += nt) Rc[l * ld + row] = Rc[(l + 1) * ld + row];
            }
            __syncthreads();
            // Givens rotations of rows (j, j+1) of R restore the triangle
            for (int j = lidx; j < nan_; ++j) {
                const T a = Rc[j * ld + j], b = Rc[j * ld + j + 1];
                const T h2 = a * a + b * b;
                const T hinv = h2 > T(0) ? rsqrt_(h2) : T(0);
                const T c_ = h2 > T(0) ? a * hinv : T(1);
                const T s_ = b * hinv;
                __syncthreads();
                if (tid == 0) {
                    cs[j] = c_;
                    sn[j] = s_;
                }
                for (int l = j + tid; l < nan_; l += nt) {
                    const T u = Rc[l * ld + j], v = Rc[l * ld + j + 1];
                    Rc[l * ld + j] = c_ * u + s_ * v;
                    Rc[l * ld + j + 1] = c_ * v - s_ * u;
                }
                __syncthreads();
            }
            // the same rotations on columns (j, j+1) of J and M, rows are thread-private
            for (int task = tid; task < n + m; task += nt) {
                T *row = task < n ? J + task * ld : M + (task - n) * ld;
                T u = row[lidx];
                for (int j = lidx; j < nan_; ++j) {
                    const T v = row[j + 1];
                    row[j] = cs[j] * u + sn[j] * v;
                    u = cs[j] * v - sn[j] * u;
                }
                row[nan_] = u;
            }
            // R^-1 of the reduced factor: thread j solves R y = e_j into column j
            for (int j = tid; j < nan_; j += nt) {
                Ri[j * ld + j] = T(1) / Rc[j * ld + j];
                for (int i = j - 1; i >= 0; --i) {
                    T s = T(0);
                    for (int k = i + 1; k <= j; ++k) s += Rc[k * ld + i] * Ri[j * ld + k];
                    Ri[j * ld + i] = -s / Rc[i * ld + i];
                }
            }
            na = nan_;
            cont = true;
            __syncthreads();
        }
    }

    // ---- outputs  // @phase CTA outputs
    __syncthreads();
    {
        // non-finite data: report a numerical failure instead of NaN inputs with status 0
        int bad = 0;
        for (int l = tid; l < n; l += nt) bad |= !(abs_(x[l]) < Num<T>::inf());
        if (st == 0 && __syncthreads_or(bad)) st = 3;
    }
    for (int l = tid; l < n; l += nt) {
        const T xo = (st == 0) ? x[l] : Num<T>::nan();
        if (p.U) static_cast<T *>(p.U)[(size_t)inst * n + l] = xo;
        for (int r = 0; r < p.npeers; ++r) static_cast<T *>(p.peerU[r])[(size_t)(p.row_off + inst) * n + l] = xo;
    }
    if (tid == 0) {
        if (p.status) p.status[inst] = st;
        if (p.iters) p.iters[inst] = it;
        for (int r = 0; r < p.npeers; ++r)
            if (p.peer_status[r]) p.peer_status[r][p.row_off + inst] = st;
    }
    if (p.Z) {
        T *Zb = static_cast<T *>(p.Z) + (size_t)inst * m;
        for (int r = tid; r < m; r += nt) Zb[r] = T(0);
        __syncthreads();
        if (st == 0)
            for (int l = tid; l < na; l += nt) Zb[aidx[l]] = lam[l];
    }
    __syncthreads();  // the next instance of this CTA reuses every buffer
    }  // instances
}

// Condense-only CTA kernel (MPCQP fields for n > 32).
template <typename T>
__global__ void __launch_bounds__(256) mpc_condense_cta_kernel(const SolveParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *sm = reinterpret_cast<T *>(smem_raw);
    const int tid = threadIdx.x, nt = blockDim.x;
    const int n = p.n, m = p.m;
    const CtaLay L = cta_layout(n, m, p.nx, (int)sizeof(T));
    T *mat, *vec;
    cta_bases<T>(p, L, sm, mat, vec);
    for (long long inst = blockIdx.x; inst < p.batch; inst += gridDim.x) {
        cta_condense<T, true>(p, L, mat, vec, inst);
        const T *M = mat + L.oM, *PL = mat + L.oPL;
        if (p.P)
            for (int idx = tid; idx < n * n; idx += nt)
                static_cast<T *>(p.P)[(size_t)inst * n * n + idx] = PL[(idx / n) * L.ld + idx % n];
        if (p.q)
            for (int c = tid; c < n; c += nt) static_cast<T *>(p.q)[(size_t)inst * n + c] = vec[L.o_q + c];
        if (p.G)
            for (int idx = tid; idx < m * n; idx += nt)
                static_cast<T *>(p.G)[(size_t)inst * m * n + idx] = M[(idx / n) * L.ld + idx % n];
        if (p.h)
            for (int r = tid; r < m; r += nt) static_cast<T *>(p.h)[(size_t)inst * m + r] = vec[L.o_hs + r];
        __syncthreads();
    }
}

}  // namespace qpmpc
