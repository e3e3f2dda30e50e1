// mpc_host_params.h -- host side of the C ABI that does not touch the CUDA
// runtime: argument checking, the (NP, MR) variant table and the translation of
// a qpmpc_b200_desc into the kernels' SolveParams.  Shared by qpmpc_b200.cu and
// by the host emulator build (tests/emu/), so that the CPU test-suite drives the
// device source with exactly the parameters a launch would get.
#pragma once

#include <cstring>

#include "../../include/qpmpc_b200.h"
#include "mpc_common.cuh"

namespace qpmpc {

struct Variant {
    int np, mr;
    bool mreg;
    bool paired;
};

// Smallest compiled (NP, MR) that holds n variables and m constraint rows.  With `paired` rows
// (desc.paired: every C_k, D_k is [M; -M]) one stored row per lane covers m = 2 NP rows.
inline bool pick_variant(int n, int m, Variant *out, bool paired = false) {
    static const Variant table[] = {{8, 2, true, false},   {8, 4, true, false},   {16, 2, true, false},
                                    {16, 4, false, false}, {32, 2, false, false}, {32, 4, false, false}};
    if (paired)
        for (int np = 8; np <= 32; np *= 2)
            if (n <= np && m <= 2 * np) {
                *out = Variant{np, 1, true, true};
                return true;
            }
    for (const Variant &v : table)
        if (n <= v.np && m <= v.np * v.mr) {
            *out = v;
            return true;
        }
    return false;
}

inline int check_desc(const qpmpc_b200_desc *d, const qpmpc_b200_operands *in) {
    if (!d || !in) return QPMPC_B200_EINVAL;
    if (d->batch < 0 || d->N <= 0 || d->nx <= 0 || d->nu <= 0 || d->nc < 0) return QPMPC_B200_EINVAL;
    if (d->dtype != QPMPC_B200_F64 && d->dtype != QPMPC_B200_F32) return QPMPC_B200_EINVAL;
    if (!(d->w_u > 0.0)) return QPMPC_B200_EWEIGHT;           // mpc_problem.py:104-107
    if (!d->has_wt && !d->has_wx) return QPMPC_B200_EWEIGHT;  // mpc_problem.py:108-111
    auto mat_ok = [](int mode, const void *p, bool optional) {
        if (mode == QPMPC_B200_ABSENT) return optional;
        return mode >= QPMPC_B200_SHARED_LTI && mode <= QPMPC_B200_BATCH_LTV && p != nullptr;
    };
    if (!mat_ok(d->mode_A, in->A, false) || !mat_ok(d->mode_B, in->B, false)) return QPMPC_B200_EINVAL;
    if (!mat_ok(d->mode_C, in->C, true) || !mat_ok(d->mode_D, in->D, true)) return QPMPC_B200_EINVAL;
    if (d->nc > 0 && !mat_ok(d->mode_e, in->e, false)) return QPMPC_B200_EINVAL;
    if (d->mode_x0 == QPMPC_B200_VEC_ABSENT || !in->x0) return QPMPC_B200_EINVAL;  // mpc_qp.py:49-51
    if (d->mode_goal != QPMPC_B200_VEC_ABSENT && !in->goal) return QPMPC_B200_EINVAL;
    if (d->mode_targets != QPMPC_B200_VEC_ABSENT && !in->targets) return QPMPC_B200_EINVAL;
    return 0;
}

inline void set_matrix(OperandView *v, int mode, const void *ptr, int item, int N) {
    v->ptr = nullptr;
    v->sz = v->step = v->per_instance = v->smem_off = 0;
    if (mode == QPMPC_B200_ABSENT || item == 0) return;
    const bool ltv = (mode == QPMPC_B200_SHARED_LTV || mode == QPMPC_B200_BATCH_LTV);
    v->ptr = ptr;
    v->sz = item * (ltv ? N : 1);
    v->step = ltv ? item : 0;
    v->per_instance = (mode == QPMPC_B200_BATCH_LTI || mode == QPMPC_B200_BATCH_LTV);
}

inline void set_vector(OperandView *v, int mode, const void *ptr, int size) {
    v->ptr = nullptr;
    v->sz = v->step = v->per_instance = v->smem_off = 0;
    if (mode == QPMPC_B200_VEC_ABSENT) return;
    v->ptr = ptr;
    v->sz = size;
    v->per_instance = (mode == QPMPC_B200_VEC_BATCH);
}

// Fill everything of SolveParams that does not depend on the kernel variant.
inline void fill_params(const qpmpc_b200_desc *d, const qpmpc_b200_operands *in, SolveParams *p) {
    std::memset(p, 0, sizeof(*p));
    p->batch = d->batch;
    p->N = d->N;
    p->nx = d->nx;
    p->nu = d->nu;
    p->nc = d->nc;
    p->n = d->N * d->nu;
    p->m = d->N * d->nc;
    set_matrix(&p->op[OP_A], d->mode_A, in->A, d->nx * d->nx, d->N);
    set_matrix(&p->op[OP_B], d->mode_B, in->B, d->nx * d->nu, d->N);
    set_matrix(&p->op[OP_C], d->mode_C, in->C, d->nc * d->nx, d->N);
    set_matrix(&p->op[OP_D], d->mode_D, in->D, d->nc * d->nu, d->N);
    set_matrix(&p->op[OP_E], d->mode_e, in->e, d->nc, d->N);
    set_vector(&p->op[OP_X0], d->mode_x0, in->x0, d->nx);
    set_vector(&p->op[OP_GOAL], d->mode_goal, in->goal, d->nx);
    set_vector(&p->op[OP_TGT], d->mode_targets, in->targets, d->N * d->nx);
    p->has_wt = d->has_wt != 0;
    p->has_wx = d->has_wx != 0;
    p->w_t = d->has_wt ? d->w_t : 0.0;
    p->w_x = d->has_wx ? d->w_x : 0.0;
    p->w_u = d->w_u;
    // q follows update_cost_vector (mpc_qp.py:139-149): a term needs weight >
    // 1e-10 (mpc_problem.py:146,159); a missing goal aborts before the stage
    // term is reached, a missing target trajectory drops only the stage term.
    const bool t_on = d->has_wt && d->w_t > 1e-10;
    const bool x_on = d->has_wx && d->w_x > 1e-10;
    const bool have_goal = d->mode_goal != QPMPC_B200_VEC_ABSENT;
    const bool have_tgt = d->mode_targets != QPMPC_B200_VEC_ABSENT;
    p->q_wt = t_on && have_goal;
    p->q_wx = x_on && have_tgt && !(t_on && !have_goal);
    p->max_iter = d->max_iter > 0 ? d->max_iter : 10 * (p->n + p->m) + 50;
    p->tol = d->tol > 0.0 ? d->tol : 1e-9;
}

}  // namespace qpmpc
