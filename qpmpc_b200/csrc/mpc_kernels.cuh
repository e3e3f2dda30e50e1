// mpc_kernels.cuh -- fused condense + dual active-set QP kernel (sm_100a).
//
// One group of NP lanes (NP = 8, 16 or 32, the padded number of decision
// variables n = N*nu) owns one MPC instance; a warp carries 32/NP instances.
// Lane l is variable l; it owns MR constraint rows (l, l+NP, ...) and keeps
// its rows of the n-column matrices in REGISTERS.  Lanes talk through a small
// shared-memory region per instance and warp shuffles; between the staging of
// the inputs and the tail there is no block-level synchronisation.
//
// Phases (reference citations relative to /root/reference):
//   0  stage A,B,C,D,e,x0,goal,targets of the CTA's instances into shared
//      memory: one 1-D bulk TMA copy per operand (cp.async.bulk + mbarrier).
//   A  condensing, qpmpc/mpc_qp.py:53-105 and :139-149.  Lane l rolls column
//      l of psi_k and the free response phi_k x0 in registers (nx = 2, 3, 4
//      compiled; other nx go through shared memory), emits h and q_l and row l
//      of P.  Time-invariant A, B, C: G is kept as a Toeplitz table (nc x n
//      numbers) instead of the dense m x n matrix.
//   B  Cholesky P = L L' (row per lane, columns published in shared memory),
//      then lane-local forward substitutions with L: t = L^-1 q and the owned
//      rows of M = G L^-T; violations at the unconstrained optimum are -M t - h.
//   C  Goldfarb-Idnani dual active-set iteration -- the algorithm of the
//      quadprog backend behind qpsolvers.solve_problem (qpmpc/solve_mpc.py:43)
//      -- on (M, R^-1): a Householder reflection when a constraint enters,
//      Givens rotations derived from R^-1 alone when one leaves; R^-1 gains the
//      column [-r/beta; 1/beta] for free.  J = L^-T Q is never formed when M
//      lives in registers.  Exact on exit.
//   D  x = -P^-1 (q + G_A' lambda) from the multipliers (two triangular solves
//      with L), then U (coalesced), status, iterations, optional multipliers,
//      optional stores into peer buffers (fused gather).
#pragma once

#include "mpc_common.cuh"
#include "mpc_plant.cuh"
#include "mpc_factor.cuh"  // FactorLay: the record of the shared-model fast path (PRE)

// Resident CTAs of 128 threads per SM the register-resident variants are
// compiled for (caps registers per thread at 65536 / (128 * QPMPC_MINB)).
#ifndef QPMPC_MINB
#define QPMPC_MINB 4
#endif
// single precision needs half the registers for the same arrays: 4 CTAs of 8 warps per SM
// (64 registers, 136 B of spills; config 4's 8 192 instances then fit in ONE wave of the
// machine: measured 219 (3 CTAs, 80 registers) -> 250 M solves/s; 2 CTAs at 125 registers: 233)
#ifndef QPMPC_MINB_F32
#define QPMPC_MINB_F32 4
#endif
// paired-row variants (one stored row per lane): resident CTAs of 8 warps per SM, double precision
#ifndef QPMPC_MINB_PAIRED
#define QPMPC_MINB_PAIRED 2
#endif
// shared-model variant (PRE), double precision: resident CTAs of 8 warps per SM (3, i.e. 80
// registers and 224 B of spills, was measured on the fused loops: 501 -> 390, 458 -> 215 M solves/s)
#ifndef QPMPC_MINB_PRE
#define QPMPC_MINB_PRE 2
#endif
// ... and the largest CTA they are launched with (threads)
#ifndef QPMPC_THREADS_PAIRED
#define QPMPC_THREADS_PAIRED 256
#endif
#ifndef QPMPC_SYNC_TAIL
#define QPMPC_SYNC_TAIL 1
#endif


namespace qpmpc {

// PAIRED: the constraint rows come in pairs [G+; -G+] (two-sided bounds, desc.paired): only the
// MR * NP rows G+ are kept (as rows of M), each standing for both of its signs; h and a dense G
// (time-varying models) still hold all 2 MR NP rows.
// PRE: shared-model fast path (mpc_factor.cuh): L, M, G come from the record of the model, the
// per-instance region keeps only h, R^-1 and the small vectors.
template <typename T, int NP, int MR, bool MREG, bool RS = false, bool PAIRED = false, bool PRE = false>
struct Lay {
    static constexpr int MP = MR * NP;     // padded (stored) constraint rows
    static constexpr int HP = PAIRED ? 2 * MP : MP;  // padded rows of h and of the dense G
    static constexpr int LDG = HP + 1;     // G / M by columns: Gc[c*LDG + row]
    static constexpr int LDL = NP + 2;     // L by columns:     Lc[c*LDL + row]
    static constexpr int szG = ((NP * LDG + 3) / 4) * 4;
    static constexpr int oH = 0;           // hs[HP]
    static constexpr int oRL = oH + HP;    // psi exchange (A), Lc (B and, if MREG, the final solve)
    static constexpr int szRL = PRE ? 0 : ((NP * LDL + 3) / 4) * 4;
    // R^-1 by columns: its own region when L must survive (MREG), else over L
    static constexpr int oRi = MREG ? oRL + szRL : oRL;
    static constexpr int oV = oRL + szRL + (MREG ? NP * NP : 0);  // qs, xs, dv, dd, d2 [NP each], sc[8]
    static constexpr int szV = 5 * NP + 8;
    // RS: the per-row constants of the iteration (vtol, ginv, |M_i|^2, MP each) live in shared
    // memory instead of registers (an option of the NP = 16 register-resident variant)
    static constexpr bool ROWS_IN_SMEM = RS;
    static constexpr int oW = oV + szV;
    static constexpr int fixed = oW + (ROWS_IN_SMEM ? 3 * MP : 0);  // runtime-sized tail follows (TailLay)
    static_assert(PRE || NP * NP <= szRL, "R^-1 must fit in the L region");
    static_assert(PRE || 8 * NP <= szRL, "psi exchange buffers must fit in the L region");
    static_assert(!PAIRED || (MREG && !RS), "paired rows: register-resident M, row constants in registers");
    static_assert(!PRE || (PAIRED && MR == 1), "shared-model fast path: paired rows, one stored row per lane");
};

// Scratch of the generic (any nx) condensing: psi[2][nx][NP], xbar[2][nx],
// phi[2][nx][nx].
__host__ __device__ inline int generic_condense_elems(int NP, int nx) {
    return (2 * nx * NP + 2 * nx + 2 * nx * nx + 3) / 4 * 4;
}
__host__ __device__ inline bool nx_in_registers(int nx) { return nx >= 2 && nx <= 4; }

// Runtime-sized tail of the per-instance work region, after Lay::fixed:
//   gt   nc x n   Toeplitz table of G (time-invariant A, B, C only), see g_toeplitz()
//   G    szG      dense G by columns; absent when the table replaces it and M
//                 lives in registers
//   scr           scratch of the generic condensing (nx outside 2..4)
struct TailLay {
    int gt_off, g_off, scr_off, total;  // offsets from the start of the instance region
};
__host__ __device__ inline TailLay tail_layout(int fixed, int szG, int NP, int nx, int nc, int n, bool toeplitz,
                                               bool mreg) {
    TailLay t;
    int o = fixed;
    t.gt_off = o;
    o += toeplitz ? (nc * n + 3) / 4 * 4 : 0;
    t.g_off = o;
    o += (toeplitz && mreg) ? 0 : szG;
    t.scr_off = o;
    o += nx_in_registers(nx) ? 0 : generic_condense_elems(NP, nx);
    t.total = o;
    return t;
}

// Entry (row, c) of G from the Toeplitz table.  With time-invariant A, B, C
// the block G[(k, r), c] = C_r A^(k-1-j) B[:, jj] (c = j nu + jj, j < k) depends
// on k nu - 1 - c only: gt[r*n + e] = C_r psi_N[:, n-1-e] holds all of them.
// Block column k carries D_k (mpc_qp.py:73-78); everything to its right is 0.
template <typename T>
__device__ __forceinline__ T g_toeplitz(const T *gt, const T *Dk, int n, int nu, int k, int r, int c) {
    const int kb = k * nu;
    if (c < kb) return gt ? gt[r * n + kb - 1 - c] : T(0);  // no C: only D contributes
    if (Dk && c - kb < nu) return Dk[r * nu + c - kb];
    return T(0);
}

// ---------------------------------------------------------------------------
// Phase 0: stage the operands of instances [inst0, inst0 + cnt) into `inbase`.
// ---------------------------------------------------------------------------
template <typename T>  // @phase 0 stage inputs
__device__ __forceinline__ void stage_inputs(const SolveParams &p, T *inbase, int inst0, int cnt,
                                             uint64_t *bar) {
    // One thread does the address arithmetic and issues the bulk copies; it
    // leaves the set of operands it could copy next to the barrier.
    unsigned *mask_slot = reinterpret_cast<unsigned *>(bar + 1);
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        fence_barrier_init();
        unsigned tx = 0, tma_mask = 0;
#pragma unroll
        for (int o = 0; o < OP_COUNT; ++o) {
            const OperandView &v = p.op[o];
            if (v.ptr == nullptr) continue;
            const T *src = static_cast<const T *>(v.ptr) + (v.per_instance ? (size_t)inst0 * v.sz : 0);
            const unsigned bytes = (unsigned)((v.per_instance ? cnt : 1) * v.sz * (int)sizeof(T));
            if (((reinterpret_cast<uintptr_t>(src) & 15) == 0) && ((bytes & 15) == 0) && bytes > 0) {
                tx += bytes;
                tma_mask |= 1u << o;
            }
        }
        mbar_arrive_expect_tx(bar, tx);
#pragma unroll
        for (int o = 0; o < OP_COUNT; ++o) {
            if (!(tma_mask >> o & 1)) continue;
            const OperandView &v = p.op[o];
            const T *src = static_cast<const T *>(v.ptr) + (v.per_instance ? (size_t)inst0 * v.sz : 0);
            const unsigned bytes = (unsigned)((v.per_instance ? cnt : 1) * v.sz * (int)sizeof(T));
            bulk_g2s(inbase + v.smem_off, src, bytes, bar);
        }
        *mask_slot = tma_mask;
    }
    __syncthreads();
    const unsigned tma_mask = *mask_slot;
    if (tma_mask != (unsigned)p.present_mask) {
        // Operands a bulk copy cannot take (odd tail counts, unaligned views).
#pragma unroll
        for (int o = 0; o < OP_COUNT; ++o) {
            const OperandView &v = p.op[o];
            if (v.ptr == nullptr || (tma_mask >> o & 1)) continue;
            const T *src = static_cast<const T *>(v.ptr) + (v.per_instance ? (size_t)inst0 * v.sz : 0);
            const int count = (v.per_instance ? cnt : 1) * v.sz;
            for (int i = threadIdx.x; i < count; i += blockDim.x) inbase[v.smem_off + i] = src[i];
        }
        __syncthreads();
    }
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
                                             T *gt, T *hs, T *xch, T *tscr, int l, T (&Prow)[NP], T &qj,
                                             long long inst, bool valid) {
    using L = Lay<T, NP, MR, false>;  // only LDG is used here
    const int nu = p.nu, nc = p.nc, N = p.N, n = p.n;
    const bool toep = p.toeplitz != 0;
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
    const int stepA = p.op[OP_A].step, stepB = p.op[OP_B].step, stepC = p.op[OP_C].step;
    const int stepD = p.op[OP_D].step, stepE = p.op[OP_E].step;
    const T *Ak = in[OP_A], *Bk = in[OP_B], *Ck = in[OP_C], *Dk = in[OP_D], *ek = in[OP_E];
    const bool hasC = Ck != nullptr, hasD = Dk != nullptr;
    const bool needG = !toep;  // with the Toeplitz table, G rows are rebuilt from psi_N
    const int kb = l / nu, jb = l - kb * nu;  // block and position inside it of this lane's variable
    T *gcol = Gc + l * L::LDG;
    int row = 0;
#pragma unroll
    for (int t = 0; t < NX * NX; ++t) Ar[t] = Ak[t];
    // Fast path (every BASELINE config): A, B, C are time invariant (Toeplitz
    // mode) and there are two inequality rows per step.  A, B, C sit in
    // registers; a step is the two recursions, the two h rows and, with a stage
    // cost, this lane's share of q.  The stage-cost Hessian is added after the
    // loop from prefix sums over the Toeplitz columns (nu = 1 only).
    const bool stageP = p.has_wx != 0, stageq = p.q_wx != 0;
    const bool fast = !DUMP && toep && nc == 2 && (!stageP || nu == 1);
    if (fast) {
        // No N-step loop: psi_N[:, l] = A^(N-1-kb) B[:, jb] and the free response of step k,
        // y_k = A^k x0, are powers of one matrix, so lane l builds "its" two vectors from the
        // binary digits of its exponents while A is squared log2(N) times (the same squarings
        // for every lane).  Lane k < N then owns the h rows of step k; x_N = A y_(N-1).
        T Cr[2 * NX], Ap[NX * NX], y[NX];
        const int ev = (l < n) ? N - 1 - kb : 0;
#pragma unroll
        for (int t = 0; t < NX; ++t) {
            psi[t] = (l < n) ? Bk[t * nu + jb] : T(0);
            y[t] = xb[t];
            Cr[t] = hasC ? Ck[t] : T(0);
            Cr[NX + t] = hasC ? Ck[NX + t] : T(0);
        }
#pragma unroll
        for (int t = 0; t < NX * NX; ++t) Ap[t] = Ar[t];
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
            const T *ekk = ek + l * stepE;
            T h0 = ekk[0], h1 = ekk[1];
#pragma unroll
            for (int t = 0; t < NX; ++t) {
                h0 -= Cr[t] * y[t];
                h1 -= Cr[NX + t] * y[t];
            }
            hs[2 * l] = h0;
            hs[2 * l + 1] = h1;
        }
        {
            T yl[NX];
#pragma unroll
            for (int t = 0; t < NX; ++t) yl[t] = __shfl_sync(FULL_MASK, y[t], N - 1, NP);
#pragma unroll
            for (int t = 0; t < NX; ++t) {
                T b = T(0);
#pragma unroll
                for (int s = 0; s < NX; ++s) b += Ar[t * NX + s] * yl[s];
                xb[t] = b;
            }
        }
        if (stageq && l < N) {
            // y_k - target_k for the stage cost's share of q (below, once the table of psi is shared)
            const T *tg = in[OP_TGT] + l * NX;
#pragma unroll
            for (int t = 0; t < NX; ++t) xch[l * NX + t] = y[t] - tg[t];
        }
        if (stageP) {
            // P += w_x sum_k psi_k' psi_k with psi_k[:, c] = v_(k-1-c), v_e = A^e B = psi_N[:, n-1-e]:
            //   P[i][j] += w_x T_|i-j|[N-2-max(i,j)],  T_d[u] = sum_{a<=u} v_a . v_(a+d).
            // Lane d runs the prefix sum T_d; the table goes through `tscr` (N x NP).
            T *W = xch + NX * NP;  // v_e by rows of t: W[t*NP + e]
            if (l < n) {
#pragma unroll
                for (int t = 0; t < NX; ++t) W[t * NP + (n - 1 - l)] = psi[t];
            }
            __syncwarp();
            if (stageq) {
                // q_l += w_x sum_{k > l} psi_k[:, l]'(y_k - target_k),  psi_k[:, l] = v_(k-1-l)   (nu = 1)
                T acc = T(0);
                for (int k = l + 1; k < N; ++k) {
#pragma unroll
                    for (int t = 0; t < NX; ++t) acc += W[t * NP + k - 1 - l] * xch[k * NX + t];
                }
                qj += w_x * acc;
            }
            if (l < N) {
                T acc = T(0);
                for (int u = 0; u + l < N; ++u) {
#pragma unroll
                    for (int t = 0; t < NX; ++t) acc += W[t * NP + u] * W[t * NP + u + l];
                    tscr[l * NP + u] = acc;
                }
            }
            __syncwarp();
#pragma unroll
            for (int j = 0; j < NP; ++j) {
                const int hi = l > j ? l : j, d = l > j ? l - j : j - l;
                if (hi <= N - 2 && l < n && j < n) Prow[j] += w_x * tscr[d * NP + (N - 2 - hi)];
            }
            __syncwarp();
        }
    }
    for (int k = fast ? N : 0; k < N; ++k) {
        if (stepA != 0 && k > 0) {
#pragma unroll
            for (int t = 0; t < NX * NX; ++t) Ar[t] = Ak[t];
        }
        const bool inblk = (k == kb);
        // G_k = C_k psi_k + [0 .. D_k .. 0], h_k = e_k - C_k (phi_k x0)   (mpc_qp.py:67-78)
        for (int r = 0; r < nc; ++r, ++row) {
            T g = T(0), hv = ek[r];
            if (hasC) {
#pragma unroll
                for (int t = 0; t < NX; ++t) {
                    const T c = Ck[r * NX + t];
                    if (needG) g += c * psi[t];
                    hv -= c * xb[t];
                }
            }
            if (needG) {
                if (hasD && inblk) g += Dk[r * nu + jb];
                gcol[row] = g;
            }
            if (l == 0) hs[row] = hv;
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
            if (inblk) a = Bk[t * nu + jb];
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
        Ak += stepA;
        Bk += stepB;
        Ck += hasC ? stepC : 0;
        Dk += hasD ? stepD : 0;
        ek += stepE;
    }
    // Toeplitz table of G: gt[r][n-1-l] = C_r psi_N[:, l]
    if (toep && in[OP_C] && l < n) {
        for (int r = 0; r < nc; ++r) {
            T g = T(0);
#pragma unroll
            for (int t = 0; t < NX; ++t) g += in[OP_C][r * NX + t] * psi[t];
            gt[r * n + n - 1 - l] = g;
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
    using L = Lay<T, NP, MR, false>;  // only LDG is used here
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
                                                  T *gt, T *hs, T *xch, T *tscr, T *scr, int l, T (&Prow)[NP],
                                                  T &qj, long long inst, bool valid) {
    switch (p.nx) {
        case 2: condense_reg<T, NP, MR, 2, DUMP>(p, in, Gc, gt, hs, xch, tscr, l, Prow, qj, inst, valid); break;
        case 3: condense_reg<T, NP, MR, 3, DUMP>(p, in, Gc, gt, hs, xch, tscr, l, Prow, qj, inst, valid); break;
        case 4: condense_reg<T, NP, MR, 4, DUMP>(p, in, Gc, gt, hs, xch, tscr, l, Prow, qj, inst, valid); break;
        default: condense_generic<T, NP, MR, DUMP>(p, in, Gc, hs, scr, l, Prow, qj, inst, valid); break;
    }
}

// ---------------------------------------------------------------------------
// Forward substitution L y = rhs, in place, for up to three right-hand sides
// held in registers (NR of a, b, c are used); L by columns in shared memory
// (Lc), 1/L_kk in dv.  Lane-local: every lane solves its own systems and all
// lanes read the same L entries (broadcast).
// ---------------------------------------------------------------------------
template <typename T, int NP, int LDL, int NR>
__device__ __forceinline__ void fsolve(const T *Lc, const T *dv, T (&a)[NP], T (&b)[NP], T (&c)[NP]) {
    using T2 = typename Pair<T>::type;
#pragma unroll
    for (int k = 0; k < NP; ++k) {
        const T dk = dv[k];
        const T ak = a[k] * dk, bk = (NR > 1) ? b[k] * dk : T(0), ck = (NR > 2) ? c[k] * dk : T(0);
        a[k] = ak;
        if (NR > 1) b[k] = bk;
        if (NR > 2) c[k] = ck;
        if (((k + 1) & 1) && k + 1 < NP) {
            const T lv = Lc[k * LDL + k + 1];
            a[k + 1] -= lv * ak;
            if (NR > 1) b[k + 1] -= lv * bk;
            if (NR > 2) c[k + 1] -= lv * ck;
        }
#pragma unroll
        for (int j = (k + 2) & ~1; j < NP; j += 2) {
            const T2 lv = *reinterpret_cast<const T2 *>(Lc + k * LDL + j);
            a[j] -= lv.x * ak;
            a[j + 1] -= lv.y * ak;
            if (NR > 1) {
                b[j] -= lv.x * bk;
                b[j + 1] -= lv.y * bk;
            }
            if (NR > 2) {
                c[j] -= lv.x * ck;
                c[j + 1] -= lv.y * ck;
            }
        }
    }
}

// Minimum / maximum of non-negative values over the NP lanes of an instance.  NP = 32: REDUX.
// Narrower groups (several instances per warp, each with its own member mask): shuffle
// butterflies.  REDUX with one member mask per group (QPMPC_SEG_REDUX=1) gives the same results
// but the groups of a warp are then served one after the other: measured 147 -> 136 M solves/s
// on config 2, 522 -> 431 at N = 8.
#ifndef QPMPC_SEG_REDUX
#define QPMPC_SEG_REDUX 0
#endif
template <typename T, int NP>
__device__ __forceinline__ T group_min_pos(T v, unsigned segmask) {
    if (NP == 32 || QPMPC_SEG_REDUX) return seg_min_pos(v, segmask);
#pragma unroll
    for (int off = NP / 2; off > 0; off >>= 1) v = fmin(v, __shfl_xor_sync(FULL_MASK, v, off, NP));
    return v;
}
template <typename T, int NP>
__device__ __forceinline__ T group_max_pos(T v, unsigned segmask) {
    if (NP == 32 || QPMPC_SEG_REDUX) return seg_max_pos(v, segmask);
#pragma unroll
    for (int off = NP / 2; off > 0; off >>= 1) v = fmax(v, __shfl_xor_sync(FULL_MASK, v, off, NP));
    return v;
}

// ---------------------------------------------------------------------------
// The fused kernel.
//   MREG = true  (NP <= 16): the owned rows of M = G J live in registers and J
//     itself is never formed: the iteration only needs rows of M (d = -M_p,
//     G z = M2 d2), and x is recovered once, at the end, from the multipliers:
//     x = -P^-1 (q + G_A' lambda) through two triangular solves with L.
//   MREG = false (NP = 32, or 4 rows per lane at NP = 16): M overwrites G in
//     shared memory, row l of J is kept in registers and x moves with every
//     step (x += t J2 d2), as in the textbook method.
// ---------------------------------------------------------------------------
template <typename T, int NP, int MR, bool MREG, bool RS, bool PAIRED = false, bool PRE = false>  // @phase kernel prologue
__global__ void __launch_bounds__((PAIRED && !PRE && NP <= 16 && sizeof(T) == 8) ? QPMPC_THREADS_PAIRED : 256,
                                  (NP <= 16 && MREG)
                                      ? (sizeof(T) == 4 ? QPMPC_MINB_F32
                                                        : (PRE ? QPMPC_MINB_PRE : PAIRED ? QPMPC_MINB_PAIRED : QPMPC_MINB / 2))
                                      : 1)
    mpc_solve_kernel(const SolveParams p) {
    using L = Lay<T, NP, MR, MREG, RS, PAIRED, PRE>;
    using T2 = typename Pair<T>::type;
    constexpr bool HASJ = !MREG;
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
    const unsigned segmask = (NP == 32) ? FULL_MASK : (((1u << (NP & 31)) - 1u) << seg_shift);
    const int iic = warp * IPW + sub;  // instance slot inside the CTA
    const int inst0 = blockIdx.x * ipc;
    const int cnt = min(ipc, p.batch - inst0);
    const long long inst = (long long)inst0 + iic;
    const bool valid = iic < cnt;
    const int n = p.n, m = p.m;

    T *inbase = work + (size_t)ipc * p.inst_stride;
    stage_inputs<T>(p, inbase, inst0, cnt, bar);

    T *wk = work + (size_t)iic * p.inst_stride;
    T *Gc = wk + p.g_off;
    T *gt = p.op[OP_C].ptr ? wk + p.gt_off : nullptr;
    T *hs = wk + L::oH;
    T *Lc = wk + L::oRL;
    T *Ri = wk + L::oRi;  // R^-1 by columns, Ri[k*NP + row]
    T *qs = wk + L::oV;
    T *xs = qs + NP;
    T *dv = qs + 2 * NP;
    T *dd = qs + 3 * NP;
    T *d2 = qs + 4 * NP;
    T *sc = qs + 5 * NP;
    const bool toep = p.toeplitz != 0;

    const T *in[OP_COUNT];
#pragma unroll
    for (int o = 0; o < OP_COUNT; ++o) {
        const OperandView &v = p.op[o];
        // Invalid tail slots read instance 0 of the CTA: defined data, results discarded.  (In a
        // fused loop the slots are rewritten every cycle: there they keep their own, unused one.)
        const bool own = valid || (PRE && p.loop.kind != 0);
        in[o] = v.ptr ? inbase + v.smem_off + (v.per_instance ? (own ? iic : 0) * v.sz : 0) : nullptr;
    }

    // ---- shared-model fast path: the record of the model, once per CTA -------  // @phase PRE record
    const FactorLay F = factor_layout(NP, p.nx, p.N, p.q_wx != 0);
    const T *rec = inbase + p.input_elems;
    if (PRE) {
        const T *src = static_cast<const T *>(p.record);
        T *dst = inbase + p.input_elems;
        for (int i = threadIdx.x; i < F.total / 2; i += blockDim.x)
            reinterpret_cast<T2 *>(dst)[i] = reinterpret_cast<const T2 *>(src)[i];
        __syncthreads();
    }

    // ---- fused closed loop (shared-model kernel only): every cycle runs the code below on the
    // inputs staged in shared memory, then the plant moves and rewrites them in place
    const int ncyc = (PRE && p.loop.kind != 0) ? p.loop.cycles : 1;
    const bool looping = PRE && p.loop.kind != 0;
    // loop-carried state of the walking pattern (uniform over the lanes of an instance)
    int wk_index = 0, wk_sidx = 0;
    T wk_foot = T(0), wk_s0 = T(0), wk_s1 = T(0), pend_v = T(0);
    if (PRE && looping && valid) {
        if (p.loop.kind == 1) {
            pend_v = static_cast<const T *>(p.loop.v_target)[inst];
        } else {
            wk_index = p.loop.phase_index[inst];
            wk_sidx = p.loop.stride_index[inst];
            wk_foot = static_cast<const T *>(p.loop.support_foot)[inst];
            wk_s0 = static_cast<const T *>(p.loop.strides)[2 * inst];
            wk_s1 = static_cast<const T *>(p.loop.strides)[2 * inst + 1];
        }
    }
#pragma unroll 1
    for (int cyc = 0; cyc < ncyc; ++cyc) {
    // ---- phase A -----------------------------------------------------------
    T Prow[PRE ? 1 : NP];
    T qj;
    bool spd = true;
    if constexpr (!PRE) {
    // prefix-sum table of the stage-cost Hessian: any NP x NP scratch that is free before phase B
    // (the condensing code sizes G by its second integer parameter: all 2 MR NP rows when paired)
    condense_dispatch<T, NP, PAIRED ? 2 * MR : MR, false>(p, in, Gc, gt, hs, Lc, MREG ? Ri : Gc, wk + p.scr_off, l,
                                                          Prow, qj, inst, valid);
    } else {
        // q = Fx x0 - Fg goal - Ft targets (update_cost_vector, mpc_qp.py:139-149, as linear maps)
        T a0 = T(0), a1 = T(0);
        for (int t = 0; t < p.nx; ++t) {
            a0 += rec[F.oFx + t * NP + l] * in[OP_X0][t];
            if (p.q_wt) a1 += rec[F.oFg + t * NP + l] * in[OP_GOAL][t];
        }
        if (p.q_wx) {
            const T *tg = in[OP_TGT];
            const int cnt = p.N * p.nx;
            T b0 = T(0), b1 = T(0);
            int j = 0;
            for (; j + 1 < cnt; j += 2) {
                b0 += rec[F.oFt + j * NP + l] * tg[j];
                b1 += rec[F.oFt + (j + 1) * NP + l] * tg[j + 1];
            }
            if (j < cnt) b0 += rec[F.oFt + j * NP + l] * tg[j];
            a1 += b0 + b1;
        }
        qj = a0 - a1;
        spd = rec[F.oDv + l] > T(0);  // NaN if the model's Hessian is not positive definite
    }

    // ---- phase B: Cholesky, row l of L in Prow, columns published in Lc -----  // @phase B cholesky
    qs[l] = qj;
    // Shared vectors start finite: lanes of finished instances keep computing
    // with them (their results are multiplied by a zero step).
    dd[l] = T(0);
    d2[l] = T(0);
    if (l < 2) sc[l] = T(0);
    if constexpr (!PRE) {
#pragma unroll
    for (int c = 0; c < NP; ++c) {
        const T piv = __shfl_sync(FULL_MASK, Prow[c], c, NP);
        spd = spd && (piv > T(0));
        const T inv = frsqrt_(piv);
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
    }

    // Forward substitutions with L, lane-local.  t = L^-1 q always; row l of
    // J = L^-T (column l of L^-1) when J is kept; the owned rows of M = G J
    // (solve L m' = g').  // @phase B substitutions: t, J, M; violations
    T Jrow[HASJ ? NP : 1];
    T Mrow[MREG ? MR : 1][NP];
    T x = T(0);
    // unpaired: viol[s] = G_i x - h_i.  Paired: the stored row G+ stands for G+ x <= h+ and
    // -G+ x <= h-, i.e. |G+ x - cen| <= wid with cen = (h+ - h-)/2, wid = (h+ + h-)/2; viol[s]
    // holds G+ x and the violation of the worse sign is |viol[s] - cen[s]| - wid[s].
    T viol[MR];
    T cen[PAIRED ? MR : 1], wid[PAIRED ? MR : 1];
    const int half = PAIRED ? (p.nc >> 1) : 1;  // rows per step of G+
    T rowc[RS ? 1 : 3][RS ? 1 : MR];                  // vtol, ginv, |M_i|^2 in registers ...
    T *rowc_s = wk + L::oW;                            // ... or in shared memory [3][MP]
    auto rc_set = [&](int which, int s, T v) {
        if (RS)
            rowc_s[which * L::MP + l + s * NP] = v;
        else
            rowc[RS ? 0 : which][RS ? 0 : s] = v;
    };
    auto rc_get = [&](int which, int s) -> T {
        return RS ? rowc_s[which * L::MP + l + s * NP] : rowc[RS ? 0 : which][RS ? 0 : s];
    };
    bool rowvalid[MR];
    if constexpr (PRE) {  // @phase PRE vectors
        // t = L^-1 q by the explicit inverse of the record, h = e -/+ Hx x0 for the pair of this
        // lane's stored row (update_constraint_vector, mpc_qp.py:161-163), its row of M and the
        // violations -M t - h; nothing else is per instance
        __syncwarp();
        T t0 = T(0), t1 = T(0);
#pragma unroll
        for (int k = 0; k < NP; k += 2) {
            const T2 v = *reinterpret_cast<const T2 *>(qs + k);
            t0 += rec[F.oLinv + k * NP + l] * v.x;
            t1 += rec[F.oLinv + (k + 1) * NP + l] * v.y;
        }
        xs[l] = t0 + t1;
        __syncwarp();
        const int srow = l;
        rowvalid[0] = srow < (m >> 1);
        const int rk = rowvalid[0] ? srow / half : 0, rr = rowvalid[0] ? srow - rk * half : 0;
        T hx = T(0);
        for (int t = 0; t < p.nx; ++t) hx += rec[F.oHx + t * NP + srow] * in[OP_X0][t];
        const T *ek = in[OP_E] + rk * p.op[OP_E].step;
        const T hp = rowvalid[0] ? ek[rr] - hx : T(0), hm = rowvalid[0] ? ek[rr + half] + hx : T(0);
        cen[0] = rowvalid[0] ? T(0.5) * (hp - hm) : T(0);
        wid[0] = rowvalid[0] ? T(0.5) * (hp + hm) : T(1);
        T v0 = T(0), v1 = T(0);
#pragma unroll
        for (int c = 0; c < NP; c += 2) {
            const T2 tv = *reinterpret_cast<const T2 *>(xs + c);
            const T ma = rec[F.oM + c * NP + srow], mb = rec[F.oM + (c + 1) * NP + srow];
            Mrow[0][c] = ma;
            Mrow[0][c + 1] = mb;
            v0 += ma * tv.x;
            v1 += mb * tv.y;
        }
        viol[0] = rowvalid[0] ? -(v0 + v1) : T(0);
        rc_set(0, 0, Num<T>::viol_eps * (fmax(T(1), fmax(abs_(hp), abs_(hm))) + rec[F.oRowc + srow]));
        rc_set(1, 0, rowvalid[0] ? rec[F.oRowc + NP + srow] : T(1e30));
        rc_set(2, 0, rec[F.oRowc + 2 * NP + srow]);
    } else {
        T tq[NP];
#pragma unroll
        for (int c = 0; c < NP; c += 2) {
            const T2 v = *reinterpret_cast<const T2 *>(qs + c);
            tq[c] = v.x;
            tq[c + 1] = v.y;
        }
        // row s of G into dst; row norm and tolerance of the violation test
        auto load_row = [&](int s, T(&dst)[NP]) {
            const int srow = l + s * NP;  // stored row
            rowvalid[s] = srow < (PAIRED ? (m >> 1) : m);
            T g2 = T(0), g21 = T(0);
            // (k, r) of this row; its slice of the Toeplitz table runs backwards
            // from gtr[0] = G[row, k nu - 1] (see g_toeplitz)
            const int rpk = PAIRED ? half : p.nc;  // stored rows per step
            const int rk = rowvalid[s] ? srow / rpk : 0, rr = rowvalid[s] ? srow - rk * rpk : 0;
            const int row = PAIRED ? rk * p.nc + rr : srow;  // index of the (+) row in G and h
            const int kb = (toep && rowvalid[s]) ? rk * p.nu : 0;
            const T *gtr = gt ? gt + rr * n + kb - 1 : nullptr;
            const T *Dr = (toep && rowvalid[s] && in[OP_D]) ? in[OP_D] + rk * p.op[OP_D].step + rr * p.nu - kb : nullptr;
            const int kd = Dr ? kb + p.nu : kb;  // block column k holds D_k
#pragma unroll
            for (int k = 0; k < NP; k += 2) {
                T ga = T(0), gb = T(0);
                if (toep && NP >= 32) {
                    // (branches, not predicated loads: keeps the live range short where registers are scarce)
                    if (rowvalid[s]) {
                        const T *Dk = in[OP_D] ? in[OP_D] + rk * p.op[OP_D].step : nullptr;
                        ga = g_toeplitz<T>(gt, Dk, n, p.nu, rk, rr, k);
                        gb = g_toeplitz<T>(gt, Dk, n, p.nu, rk, rr, k + 1);
                    }
                } else if (toep) {
                    if (gtr && k < kb) ga = gtr[-k];
                    if (gtr && k + 1 < kb) gb = gtr[-(k + 1)];
                    if (k >= kb && k < kd) ga = Dr[k];
                    if (k + 1 >= kb && k + 1 < kd) gb = Dr[k + 1];
                } else if (rowvalid[s]) {
                    ga = Gc[k * L::LDG + row];
                    gb = Gc[(k + 1) * L::LDG + row];
                }
                dst[k] = ga;
                dst[k + 1] = gb;
                g2 += ga * ga;
                g21 += gb * gb;
            }
            g2 += g21;
            // eps * (max(1, |h_i|) + |G_i|)
            T hi = rowvalid[s] ? hs[row] : T(0);
            viol[s] = -hi;
            if (PAIRED) {
                const T hm = rowvalid[s] ? hs[row + half] : T(0);
                cen[PAIRED ? s : 0] = rowvalid[s] ? T(0.5) * (hi - hm) : T(0);
                wid[PAIRED ? s : 0] = rowvalid[s] ? T(0.5) * (hi + hm) : T(1);
                hi = fmax(abs_(hi), abs_(hm));
                viol[s] = T(0);
            }
            rc_set(0, s, Num<T>::viol_eps * (fmax(T(1), abs_(hi)) + sqrt_(g2)));
            rc_set(1, s, g2 > T(0) ? frsqrt_(g2) : T(1e30));
        };
        // after the solve src is row s of M: M t (= -G x) and |M_s|^2
        auto finish_row = [&](int s, const T(&src)[NP]) {
            T m0 = T(0), m1 = T(0), v0 = T(0), v1 = T(0);
#pragma unroll
            for (int c = 0; c < NP; c += 2) {
                m0 += src[c] * src[c];
                m1 += src[c + 1] * src[c + 1];
                v0 += src[c] * tq[c];
                v1 += src[c + 1] * tq[c + 1];
            }
            rc_set(2, s, m0 + m1);
            // G x - h with x = -P^-1 q = -J t: G x = -(G J) t = -M t
            viol[s] = rowvalid[s] ? viol[s] - (v0 + v1) : (PAIRED ? T(0) : T(-1));
            if (!MREG) {
                // In place: this lane is the only reader and writer of its rows of G.
                const int row = l + s * NP;
#pragma unroll
                for (int c = 0; c < NP; ++c) Gc[c * L::LDG + row] = src[c];
            }
        };
        if (MREG && MR == 1) {
            // one stored row per lane (paired rows): t rides along with it
            load_row(0, Mrow[0]);
            fsolve<T, NP, L::LDL, 2>(Lc, dv, tq, Mrow[0], Mrow[0]);
            finish_row(0, Mrow[0]);
        } else if (MREG) {
            // t rides along with the first two rows of M, solved in place in Mrow
            load_row(0, Mrow[0]);
            load_row(1, Mrow[MREG ? 1 : 0]);
            fsolve<T, NP, L::LDL, 3>(Lc, dv, tq, Mrow[0], Mrow[MREG ? 1 : 0]);
            finish_row(0, Mrow[0]);
            finish_row(1, Mrow[MREG ? 1 : 0]);
#pragma unroll
            for (int s0 = 2; s0 < MR; s0 += 2) {
                load_row(s0, Mrow[MREG ? s0 : 0]);
                load_row(s0 + 1, Mrow[MREG ? s0 + 1 : 0]);
                fsolve<T, NP, L::LDL, 2>(Lc, dv, Mrow[MREG ? s0 : 0], Mrow[MREG ? s0 + 1 : 0], tq);
                finish_row(s0, Mrow[MREG ? s0 : 0]);
                finish_row(s0 + 1, Mrow[MREG ? s0 + 1 : 0]);
            }
        } else {
            // t and row l of J (column l of L^-1: solve L y = e_l), x = -J t
            T jr[NP];
#pragma unroll
            for (int c = 0; c < NP; ++c) jr[c] = (c == l) ? T(1) : T(0);
            fsolve<T, NP, L::LDL, 2>(Lc, dv, tq, jr, jr);
            T x0 = T(0), x1 = T(0);
#pragma unroll
            for (int c = 0; c < NP; c += 2) {
                x0 -= jr[c] * tq[c];
                x1 -= jr[c + 1] * tq[c + 1];
            }
            x = x0 + x1;
#pragma unroll
            for (int c = 0; c < NP; ++c) Jrow[HASJ ? c : 0] = jr[c];
#pragma unroll
            for (int s0 = 0; s0 < MR; s0 += 2) {
                T r0[NP], r1[NP];
                load_row(s0, r0);
                load_row(s0 + 1, r1);
                fsolve<T, NP, L::LDL, 2>(Lc, dv, r0, r1, r1);
                finish_row(s0, r0);
                finish_row(s0 + 1, r1);
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
    __syncwarp();

    // ---- phase C: dual active-set iteration ---------------------------------  // @phase C select row (step 1)
    // R^-1 starts as the zero matrix: its product with the zero-padded d1 then
    // needs no guards (entries below the diagonal stay zero throughout).  When
    // J is kept, L is dead and R^-1 takes its place; otherwise L survives for
    // the final solve and R^-1 has a region of its own.
#pragma unroll
    for (int k = 0; k < NP; ++k) Ri[k * NP + l] = T(0);
    const int max_iter = p.max_iter;
    int na = 0, it = 0;
    int st = spd ? 0 : 3;
    if (PAIRED) {
        // a pair with h+ + h- < 0 admits no point at all: infeasible, as the unpaired iteration
        // finds out when the second member meets the first in the active set (t1 = t2 = inf)
        bool bad = false;
#pragma unroll
        for (int s = 0; s < MR; ++s) bad = bad || (rowvalid[s] && wid[PAIRED ? s : 0] < -rc_get(0, s));
        if ((__ballot_sync(FULL_MASK, bad) & segmask) && st == 0) st = 2;
    }
    bool done = !valid || st != 0 || m == 0;
    bool cont = false;
    int pidx = 0;
    bool pneg = false;  // paired: the candidate is the (-) member of its pair
    T lam = T(0), lamp = T(0);
    int aidx = -1;
    unsigned actbits = 0;
    const T INF = Num<T>::inf();
    __syncwarp();

    while (true) {
        {
            // step 1: most violated inactive row, relative to its norm
            const bool sel = !done && !cont;
            T best = T(0);
            int bi = 0;
#pragma unroll
            for (int s = 0; s < MR; ++s) {
                const T off = PAIRED ? viol[s] - cen[PAIRED ? s : 0] : T(0);
                const T vs = PAIRED ? abs_(off) - wid[PAIRED ? s : 0] : viol[s];
                const T score = vs * rc_get(1, s);
                if (sel && rowvalid[s] && !((actbits >> s) & 1) && vs > rc_get(0, s) && score > best) {
                    best = score;
                    bi = (l + s * NP) | ((PAIRED && off < T(0)) ? 0x8000 : 0);
                }
            }
            // single-precision keys: the ranking is a heuristic, any violated row is a valid pivot
            const float bestf = (float)best;
            const float top = group_max_pos<float, NP>(best > T(0) ? fmaxf(bestf, 1e-37f) : 0.f, segmask);
            const unsigned win = __ballot_sync(FULL_MASK, best > T(0) && fmaxf(bestf, 1e-37f) == top) & segmask;
            const int cand_p = __shfl_sync(FULL_MASK, bi, __ffs(win) - 1);
            if (sel) {
                if (!(top > 0.f)) {
                    done = true;  // primal feasible: optimal
                } else {
                    pidx = cand_p & 0x7fff;
                    pneg = (cand_p & 0x8000) != 0;
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
        // Row p of M is -d (d = J' n_p): the iteration works with m = -d directly.
        // dd holds its active part m1 (zero-padded), d2 the rest m2.  // @phase C publish d
        const int owner = pidx % NP, pslot = pidx / NP;
        T dl;
        if (MREG) {
            if (act && l == owner) {
#pragma unroll
                for (int s = 0; s < MR; ++s) {
                    if (s == pslot) {
                        // the (-) member of a pair is the row -G+: row p of M is -Mrow
                        const T sg = (PAIRED && pneg) ? T(-1) : T(1);
#pragma unroll
                        for (int c = 0; c < NP; c += 2) {
                            T2 v;
                            v.x = sg * Mrow[MREG ? s : 0][c];
                            v.y = sg * Mrow[MREG ? s : 0][c + 1];
                            *reinterpret_cast<T2 *>(dd + c) = v;
                        }
                        sc[0] = PAIRED ? abs_(viol[s] - cen[PAIRED ? s : 0]) - wid[PAIRED ? s : 0] : viol[s];
                        sc[1] = rc_get(2, s);
                    }
                }
            }
            __syncwarp();
            dl = dd[l];
        } else {
            dl = Gc[l * L::LDG + pidx];
            if (act && l == owner) {
#pragma unroll
                for (int s = 0; s < MR; ++s) {
                    if (s == pslot) {
                        sc[0] = viol[s];
                        sc[1] = rc_get(2, s);
                    }
                }
            }
        }
        dd[l] = (l < na) ? dl : T(0);
        d2[l] = (l >= na) ? dl : T(0);
        __syncwarp();
        // -G z = M2 m2 (owned rows), |m2|^2, and -z = J2 m2 when J is kept  // @phase C z, Gz
        T z = T(0), z1 = T(0), a2 = T(0), a21 = T(0);
        T gz[MR];
#pragma unroll
        for (int s = 0; s < MR; ++s) gz[s] = T(0);
#pragma unroll
        for (int c = 0; c < NP; c += 2) {
            const T2 v = *reinterpret_cast<const T2 *>(d2 + c);
            if (HASJ) {
                z += Jrow[HASJ ? c : 0] * v.x;
                z1 += Jrow[HASJ ? c + 1 : 0] * v.y;
            }
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
        // -r = R^-1 m1 (component l; zero on lanes >= na)  // @phase C r=R^-1 d
        T rv = T(0);
        {
            T rv1 = T(0);
            const int namax = __reduce_max_sync(FULL_MASK, act ? na : 0);
            for (int k = 0; k < namax; k += 2) {
                const T2 dk = *reinterpret_cast<const T2 *>(dd + k);
                rv += Ri[k * NP + l] * dk.x;
                rv1 += Ri[(k + 1) * NP + l] * dk.y;
            }
            rv += rv1;
        }
        // step lengths  // @phase C step length, move
        const T cand = (act && l < na && rv < T(0)) ? fmax(lam, T(0)) * rcp_(-rv) : INF;
        const T t1 = group_min_pos<T, NP>(cand, segmask);
        const unsigned bal = __ballot_sync(FULL_MASK, cand == t1 && cand < INF) & segmask;
        const int lidx = bal ? (__ffs(bal) - 1 - seg_shift) : 0;
        const T violp = sc[0], dn2 = sc[1];
        const bool zzero = !(a2 > Num<T>::dep_eps * dn2);
        const T ainv = frsqrt_(a2);  // 1 / |d2|
        const T t2 = zzero ? INF : violp * (ainv * ainv);
        if (act && t1 == INF && t2 == INF) {
            st = 2;  // infeasible
            done = true;
            act = false;
        }
        const T t = t2 < t1 ? t2 : t1;  // (neither is NaN here: plain select, not fmin)
        {
            // unconditional updates with a zero step where nothing moves (all
            // operands are finite on finished instances, so 0 * v adds nothing)
            const T tp = (act && !zzero) ? t : T(0);
            const T td = act ? t : T(0);
            if (HASJ) x -= tp * z;
#pragma unroll
            for (int s = 0; s < MR; ++s) viol[s] -= tp * gz[s];
            lam += td * rv;
            lamp += td;
        }
        const bool full = act && !zzero && t2 <= t1;
        const bool part = act && !full;

        if (__any_sync(FULL_MASK, full)) {  // @phase C add constraint (Householder)
            // Constraint p enters: reflect d2 = -m2 onto beta e_na.  H = I - tau v v',
            // v = d2 - beta e_na = -(m2 + beta e_na), applied to columns >= na of M (and J).
            const int nac = na < NP ? na : NP - 1;
            const T mna = d2[nac];
            const T alpha = a2 * ainv;
            // (instances of the warp that do not add a row ride along with beta = tau = 0: their
            // a2 may be zero and alpha not finite)
            const T beta = full ? ((mna < T(0)) ? -alpha : alpha) : T(0);
            const T binv = (mna < T(0)) ? -ainv : ainv;
            const T tau = full ? rcp_(a2 + beta * mna) : T(0);
            // One instance per warp (NP = 32): na is uniform, so one entry of a register row can be
            // taken by a switch and the product M v reuses the product with d2 from above.  With
            // several instances per warp the switch would diverge: two passes over d2 are cheaper.
            constexpr bool ONEPASS = (NP == 32);
            if (ONEPASS) {
                // -v = d2 + beta e_na: the products M v (and J v) are the products with d2 computed
                // above plus one entry of the row -- a switch over registers, na being uniform over
                // the lanes of an instance -- and the update is one pass over d2 plus that entry
                T dj = T(0);
                if (HASJ) dj = (z + reg_get<T, HASJ ? NP : 1>(Jrow, HASJ ? nac : 0) * beta) * tau;
                T dm[MR];
#pragma unroll
                for (int s = 0; s < MR; ++s) {
                    const T mcol = MREG ? reg_get<T, NP>(Mrow[MREG ? s : 0], nac) : Gc[nac * L::LDG + l + s * NP];
                    dm[s] = (gz[s] + mcol * beta) * tau;
                }
#pragma unroll
                for (int c = 0; c < NP; c += 2) {
                    const T2 v = *reinterpret_cast<const T2 *>(d2 + c);
                    if (HASJ) {
                        Jrow[HASJ ? c : 0] -= dj * v.x;
                        Jrow[HASJ ? c + 1 : 0] -= dj * v.y;
                    }
#pragma unroll
                    for (int s = 0; s < MR; ++s) {
                        mset(s, c, mget(s, c) - dm[s] * v.x);
                        mset(s, c + 1, mget(s, c + 1) - dm[s] * v.y);
                    }
                }
                if (HASJ) reg_sub<T, HASJ ? NP : 1>(Jrow, HASJ ? nac : 0, dj * beta);
#pragma unroll
                for (int s = 0; s < MR; ++s) {
                    if (MREG)
                        reg_sub<T, NP>(Mrow[MREG ? s : 0], nac, dm[s] * beta);
                    else
                        Gc[nac * L::LDG + l + s * NP] -= dm[s] * beta;
                }
            } else {
                __syncwarp();
                if (full && l == na) d2[l] = mna + beta;  // d2 becomes -v
                __syncwarp();
                T dj = T(0), dj1 = T(0);
                T dm[MR];
#pragma unroll
                for (int s = 0; s < MR; ++s) dm[s] = T(0);
#pragma unroll
                for (int c = 0; c < NP; c += 2) {
                    const T2 v = *reinterpret_cast<const T2 *>(d2 + c);
                    if (HASJ) {
                        dj += Jrow[HASJ ? c : 0] * v.x;
                        dj1 += Jrow[HASJ ? c + 1 : 0] * v.y;
                    }
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
                    if (HASJ) {
                        Jrow[HASJ ? c : 0] -= dj * v.x;
                        Jrow[HASJ ? c + 1 : 0] -= dj * v.y;
                    }
#pragma unroll
                    for (int s = 0; s < MR; ++s) {
                        mset(s, c, mget(s, c) - dm[s] * v.x);
                        mset(s, c + 1, mget(s, c + 1) - dm[s] * v.y);
                    }
                }
            }
            if (full) {
                // R gains the column [d1; beta]: R^-1 gains [-r / beta; 1 / beta]
                if (l < na) Ri[na * NP + l] = rv * binv;  // rv holds -r
                if (l == na) {
                    Ri[na * NP + na] = binv;
                    lam = lamp;
                    // paired: stored row | sign bit; turned into the row's index in G / h after the loop
                    aidx = PAIRED ? (pidx | (pneg ? 0x8000 : 0)) : pidx;
                }
                if (l == owner) actbits |= 1u << pslot;
                ++na;
                cont = false;
            }
        }
        if (__any_sync(FULL_MASK, part)) {  // @phase C drop constraint (Givens)
            // Constraint at active position lidx leaves; p stays the candidate.
            int cidx = __shfl_sync(FULL_MASK, aidx, lidx, NP);
            if (PAIRED) cidx &= 0x7fff;  // stored row
            if (part && l == cidx % NP) actbits &= ~(1u << (cidx / NP));
            const T lam_n = __shfl_down_sync(FULL_MASK, lam, 1, NP);
            const int aidx_n = __shfl_down_sync(FULL_MASK, aidx, 1, NP);
            const int nan_ = na - 1;
            const bool mover = part && l >= lidx && l < nan_;
            if (mover) {
                lam = lam_n;
                aidx = aidx_n;
            }
            if (part && l == nan_) lam = T(0);  // the vacated position carries no multiplier
            // Downdate through R^-1 alone: rotations of adjacent columns (j, j+1),
            // j = lidx .. na-2, that zero row lidx of R^-1 Q up to its last entry;
            // Q is the factor that re-triangularises R without column lidx, so the
            // same rotations act on columns (j, j+1) of M (and J), and R'^-1 is
            // R^-1 Q with row lidx and the last column removed.
            T a = part ? Ri[lidx * NP + lidx] : T(1);
#pragma unroll
            for (int j = 0; j < NP - 1; ++j) {
                const bool rot = part && j >= lidx && j < nan_;
                if (!__any_sync(FULL_MASK, rot)) continue;
                const T b = rot ? Ri[(j + 1) * NP + lidx] : T(0);
                __syncwarp();  // b is read before its owner rotates it
                const T h2 = a * a + b * b;
                const T hinv = h2 > T(0) ? frsqrt_(h2) : T(0);
                const T cs = h2 > T(0) ? b * hinv : T(1);
                const T sn = -a * hinv;
                if (rot) {
                    a = h2 * hinv;
                    if (l <= j + 1) {
                        // R^-1 is upper triangular: entry (j+1, j) is zero
                        const T u = (l <= j) ? Ri[j * NP + l] : T(0), v = Ri[(j + 1) * NP + l];
                        Ri[j * NP + l] = cs * u + sn * v;
                        Ri[(j + 1) * NP + l] = cs * v - sn * u;
                    }
                    if (HASJ) {
                        const T u = Jrow[HASJ ? j : 0], v = Jrow[HASJ ? j + 1 : 0];
                        Jrow[HASJ ? j : 0] = cs * u + sn * v;
                        Jrow[HASJ ? j + 1 : 0] = cs * v - sn * u;
                    }
#pragma unroll
                    for (int s = 0; s < MR; ++s) {
                        const T mu = mget(s, j), mv = mget(s, j + 1);
                        mset(s, j, cs * mu + sn * mv);
                        mset(s, j + 1, cs * mv - sn * mu);
                    }
                }
            }
            __syncwarp();
            // remove row lidx: rows above move down one position
            {
                const int kmax = __reduce_max_sync(FULL_MASK, part ? nan_ : 0);
                for (int k = 0; k < kmax; ++k) {
                    const bool mv = mover && l <= k;
                    T v = T(0);
                    if (mv) v = Ri[k * NP + l + 1];
                    __syncwarp();
                    if (mv) Ri[k * NP + l] = v;
                    if (part && l == k + 1 && l > lidx) Ri[k * NP + l] = T(0);  // keep R^-1 upper triangular
                }
            }
            if (part) {
                na = nan_;
                cont = true;
            }
        }
        __syncwarp();
    }

    const int araw = aidx;  // (paired: stored row | sign bit)
    if (PAIRED && aidx >= 0) {
        // stored row | sign -> index of the row in G / h
        const int sr = aidx & 0x7fff, sk = sr / half;
        aidx = sk * p.nc + (sr - sk * half) + ((aidx & 0x8000) ? half : 0);
    }
    // The warps of the CTA leave the iteration at different times but hold the
    // CTA's resources until the last one is done: let them run the (unrolled,
    // one-shot) recovery code together so that they share instruction fetches.
    if (!HASJ && QPMPC_SYNC_TAIL && !looping) __syncthreads();  // (a fused loop would pay the wait every cycle)
    // ---- x from the multipliers (J not kept): x = -P^-1 (q + G_A' lambda)  // @phase D x from multipliers
    if constexpr (PRE) {
        // shared model: L^-1 (q + G_A' lambda) = t + M_A' lambda with t (still in xs) and the rows of
        // M = G+ L^-T from the record, then x = -L^-T y with L^-1 by rows: two short products,
        // no substitution (the substitutions are 2 NP dependent shuffles)
        T y = xs[l];
        const int namax = __reduce_max_sync(FULL_MASK, st == 0 ? na : 0);
        for (int i = 0; i < namax; ++i) {
            const T li = __shfl_sync(FULL_MASK, lam, i, NP);
            const int ai = __shfl_sync(FULL_MASK, araw, i, NP);
            if (i < na && st == 0) y += ((ai & 0x8000) ? -li : li) * rec[F.oMT + (ai & 0x7fff) * F.LDM + l];
        }
        __syncwarp();
        xs[l] = y;
        __syncwarp();
        T xa = T(0), xb = T(0);
#pragma unroll
        for (int k = 0; k < NP; k += 2) {
            const T2 v = *reinterpret_cast<const T2 *>(xs + k);
            xa += rec[F.oLinvT + k * NP + l] * v.x;
            xb += rec[F.oLinvT + (k + 1) * NP + l] * v.y;
        }
        x = -(xa + xb);
    } else if (!HASJ) {
        // w_l = q_l + sum_i lambda_i G[a_i, l]
        T w = qs[l];
        {
            const int nc = p.nc > 0 ? p.nc : 1;
            const int ak = (aidx >= 0) ? aidx / nc : 0, ar = (aidx >= 0) ? aidx - ak * nc : 0;
            const int namax = __reduce_max_sync(FULL_MASK, st == 0 ? na : 0);
            for (int i = 0; i < namax; ++i) {
                const T li = __shfl_sync(FULL_MASK, lam, i, NP);
                const int ai = __shfl_sync(FULL_MASK, aidx, i, NP);
                const int ki = __shfl_sync(FULL_MASK, ak, i, NP), ri = __shfl_sync(FULL_MASK, ar, i, NP);
                if (i < na && st == 0) {
                    T g;
                    if (toep) {
                        const T *Dk = in[OP_D] ? in[OP_D] + ki * p.op[OP_D].step : nullptr;
                        g = g_toeplitz<T>(gt, Dk, n, p.nu, ki, ri, l);
                    } else {
                        g = (PRE ? rec + F.oG : Gc)[l * L::LDG + ai];
                    }
                    w += li * g;
                }
            }
        }
        // forward: L y = w, lane c finalises y_c and broadcasts it
        const T *Lfin = PRE ? rec + F.oL : Lc, *dvfin = PRE ? rec + F.oDv : dv;  // the model's factor
        const T dinv = dvfin[l];
        T y = T(0);
#pragma unroll
        for (int c = 0; c < NP; ++c) {
            const T yc = __shfl_sync(FULL_MASK, w * dinv, c, NP);
            if (l == c) y = yc;
            if (l > c) w -= Lfin[c * L::LDL + l] * yc;
        }
        // backward: L' x = -y
        T s2 = -y;
#pragma unroll
        for (int c = NP - 1; c >= 0; --c) {
            const T xc = __shfl_sync(FULL_MASK, s2 * dinv, c, NP);
            if (l == c) x = xc;
            if (l < c) s2 -= Lfin[l * L::LDL + c] * xc;
        }
    }

    // ---- phase D: outputs ---------------------------------------------------  // @phase D outputs
    // Non-finite data (NaN / inf operands) slip through every comparison above:
    // an instance whose solution is not finite is reported as a numerical failure.
    {
        const unsigned nonfinite = __ballot_sync(FULL_MASK, !(abs_(x) < Num<T>::inf())) & segmask;  // all lanes vote
        if (st == 0 && nonfinite) st = 3;
    }
    if (valid && cyc + 1 == ncyc) {
        const T xo = (st == 0) ? x : Num<T>::nan();
        if (l < n && p.U) static_cast<T *>(p.U)[(size_t)inst * n + l] = xo;
        if (l == 0) {
            if (p.status) p.status[inst] = st;
            if (p.iters) p.iters[inst] = it;
        }
        // fused gather: the same row goes straight into every peer's buffer
        // (peer-mapped pointers: the stores travel over NVLink)
        for (int r = 0; r < p.npeers; ++r) {
            if (l < n) static_cast<T *>(p.peerU[r])[(size_t)(p.row_off + inst) * n + l] = xo;
            if (l == 0 && p.peer_status[r]) p.peer_status[r][p.row_off + inst] = st;
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
    if constexpr (PRE) {
        if (looping) {  // @phase PRE loop: plant and next cycle's vectors
            // first input of the plan (nu = 1: variable 0), zero when the cycle has no plan
            const T u0 = __shfl_sync(FULL_MASK, (st == 0) ? x : T(0), 0, NP);
            T *xm = const_cast<T *>(in[OP_X0]);
            T *gm = const_cast<T *>(in[OP_GOAL]);
            const LoopDev &lp = p.loop;
            const int nxs = p.nx;
            __syncwarp();
            if (valid && l == 0) {
                if (st != 0 && lp.unsolved) atomicAdd(lp.unsolved, 1);
                if (lp.iter_sum) atomicAdd(reinterpret_cast<unsigned long long *>(lp.iter_sum + cyc), (unsigned long long)it);
            }
            if (lp.kind == 1) {
                if (valid && l == 0) {
                    T r = xm[0], th = xm[1], rd = xm[2], thd = xm[3];
                    // the MPC of this cycle was solved from the state before the plant moves
                    if (lp.upright && abs_(th) <= T(1.2)) atomicAdd(lp.upright, 1);
                    pendulum_integrate<T>(r, th, rd, thd, u0, lp.substeps, (T)lp.dt, (T)lp.omega2, (T)lp.g);
                    xm[0] = r, xm[1] = th, xm[2] = rd, xm[3] = thd;
                }
                __syncwarp();
                if (valid) {
                    T *tg = const_cast<T *>(in[OP_TGT]);
                    const T r = xm[0];
                    for (int k = l; k < p.N; k += NP) pendulum_target<T>(tg + k * 4, r, pend_v, k, (T)lp.T);
                    if (l == 0) pendulum_target<T>(gm, r, pend_v, p.N, (T)lp.T);
                }
            } else {
                if (valid && l == 0) {
                    T pos = xm[0], vel = xm[1], acc = xm[2];
                    lipm_integrate<T>(pos, vel, acc, u0, lp.substeps, (T)lp.dt);
                    xm[0] = pos, xm[1] = vel, xm[2] = acc;
                }
                // PhaseStepper.advance and the foot switch (:331-334), every lane for itself
                wk_index = wk_index + 1 >= lp.nb_dsp + lp.nb_ssp ? 0 : wk_index + 1;
                if (wk_index == 0) {
                    wk_foot = wk_foot + (wk_sidx ? wk_s1 : wk_s0);
                    wk_sidx = (wk_sidx + 1) % 2;
                }
                if (valid) {
                    const LipmPhases ph = lipm_phases(wk_index, lp.nb_dsp, lp.nb_ssp, p.N);
                    const T next = wk_foot + (wk_sidx ? wk_s1 : wk_s0);
                    const T last = next + (wk_sidx ? wk_s0 : wk_s1);
                    const T hf = T(0.5) * (T)lp.foot_size, big = (T)lp.max_zmp;
                    T *em = const_cast<T *>(in[OP_E]);
                    for (int k = l; k < p.N; k += NP) lipm_bounds<T>(ph, k, wk_foot, next, last, hf, big, em[2 * k], em[2 * k + 1]);
                    if (l == 0) gm[0] = ph.n4 > 0 ? last : next, gm[1] = T(0), gm[2] = T(0);
                }
            }
            __syncwarp();
            if (valid && lp.traj && l < nxs)
                static_cast<T *>(lp.traj)[((size_t)(cyc + 1) * p.batch + inst) * nxs + l] = xm[l];
            if (valid && cyc + 1 == ncyc) {
                // what the step kernel leaves behind after the last cycle: state and next vectors
                if (l < nxs) {
                    static_cast<T *>(lp.state)[(size_t)inst * nxs + l] = xm[l];
                    static_cast<T *>(lp.goal)[(size_t)inst * nxs + l] = gm[l];
                }
                if (lp.kind == 1) {
                    const T *tg = in[OP_TGT];
                    for (int k = l; k < p.N * nxs; k += NP) static_cast<T *>(lp.targets)[(size_t)inst * p.N * nxs + k] = tg[k];
                } else {
                    const T *em = in[OP_E];
                    for (int k = l; k < 2 * p.N; k += NP) static_cast<T *>(lp.e)[(size_t)inst * p.N * 2 + k] = em[k];
                    if (l == 0) {
                        lp.phase_index[inst] = wk_index;
                        lp.stride_index[inst] = wk_sidx;
                        static_cast<T *>(lp.support_foot)[inst] = wk_foot;
                    }
                }
            }
            __syncwarp();
        }
    }
    }  // cycles
}

// ---------------------------------------------------------------------------
// Condense-only kernel: materialises the MPCQP fields for parity checks and
// for the MPCQP host class (qpmpc/mpc_qp.py:28-37).  Same phase-A code as the
// fused kernel.
// ---------------------------------------------------------------------------
// DUMP = false runs exactly the phase-A code of the fused kernel (fast paths
// included) for P, q, G, h; DUMP = true adds the Phi / Psi stacks.
template <typename T, int NP, int MR, bool DUMP>  // @phase condense-only kernel
__global__ void __launch_bounds__(128) mpc_condense_kernel(const SolveParams p) {
    using L = Lay<T, NP, MR, false>;
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
    T *gt = p.op[OP_C].ptr ? wk + p.gt_off : nullptr;
    T *hs = wk + L::oH;
    T *xch = wk + L::oRL;
    const T *in[OP_COUNT];
#pragma unroll
    for (int o = 0; o < OP_COUNT; ++o) {
        const OperandView &v = p.op[o];
        in[o] = v.ptr ? inbase + v.smem_off + (v.per_instance ? (valid ? iic : 0) * v.sz : 0) : nullptr;
    }
    T Prow[NP];
    T qj;
    condense_dispatch<T, NP, MR, DUMP>(p, in, Gc, gt, hs, xch, Gc, wk + p.scr_off, l, Prow, qj, inst, valid);
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
        if (p.toeplitz) {
            // the fused kernel's own view of G: rebuilt from the Toeplitz table
            for (int r = 0; r < m; ++r) {
                const int rk = r / p.nc, rr = r - rk * p.nc;
                const T *Dk = in[OP_D] ? in[OP_D] + rk * p.op[OP_D].step : nullptr;
                Gb[(size_t)r * n + l] = g_toeplitz<T>(gt, Dk, n, p.nu, rk, rr, l);
            }
        } else {
            for (int r = 0; r < m; ++r) Gb[(size_t)r * n + l] = Gc[l * L::LDG + r];
        }
    }
    if (p.h) {
        T *hb = static_cast<T *>(p.h) + (size_t)inst * m;
        for (int r = l; r < m; r += NP) hb[r] = hs[r];
    }
}

// ---------------------------------------------------------------------------
// Host side of the layout (launch wrappers in mpc_launch.cuh, and the host
// emulator build in tests/emu/).
// ---------------------------------------------------------------------------
// Shared-memory geometry: [16 B mbarrier][ipc work regions][CTA input region].
template <typename T>
size_t layout_smem(SolveParams *p, int fixed_elems, int szG, int np, int ipc, bool mreg, bool dense_g = false) {
    const bool lti = p->op[OP_A].step == 0 && p->op[OP_B].step == 0 && (!p->op[OP_C].ptr || p->op[OP_C].step == 0);
    p->toeplitz =
        (!dense_g && lti && nx_in_registers(p->nx) && p->nc > 0 && env_int("QPMPC_B200_NO_TOEPLITZ", 0) == 0) ? 1 : 0;
    const TailLay t = tail_layout(fixed_elems, szG, np, p->nx, p->nc, p->n, p->toeplitz != 0, mreg);
    p->gt_off = t.gt_off;
    p->g_off = t.g_off;
    p->scr_off = t.scr_off;
    p->inst_stride = t.total;
    int off = 0;
    p->present_mask = 0;
    for (int o = 0; o < OP_COUNT; ++o) {
        OperandView &v = p->op[o];
        if (!v.ptr) continue;
        p->present_mask |= 1 << o;
        v.smem_off = off;
        int elems = v.sz * (v.per_instance ? ipc : 1);
        off += (elems + 3) / 4 * 4;  // keep every region 16-byte aligned
    }
    p->input_elems = off;
    return 16 + ((size_t)ipc * p->inst_stride + off) * sizeof(T);
}

}  // namespace qpmpc
