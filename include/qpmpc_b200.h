/*
 * qpmpc_b200.h -- C ABI of the B200 batched linear-MPC engine.
 *
 * This is the drop-in boundary for ONE path of stephane-caron/qpmpc: condense
 * an MPC problem into a dense QP and solve it.  Citations are relative to the
 * reference tree (/root/reference):
 *
 *   qpmpc_b200_solve / _solve_host   replace   qpmpc/solve_mpc.py:42-43
 *       MPCQP(problem)                      (qpmpc/mpc_qp.py:39-122)
 *       qpsolvers.solve_problem(P,q,G,h)    (qpmpc/solve_mpc.py:43)
 *     and hand back what Plan reads (qpmpc/plan.py:36-39): found + x.
 *   qpmpc_b200_condense              replaces  qpmpc/mpc_qp.py:53-122,139-163
 *     (fields P, q, G, h, Phi, Psi, phi_last, psi_last of MPCQP).
 *   qpmpc_b200_integrate             replaces  qpmpc/mpc_problem.py:316-335
 *     (MPCProblem.integrate, what Plan.states calls, qpmpc/plan.py:100-109).
 *
 * Conventions: plain pointers and sizes, no C++/torch types.  Device entry
 * points take DEVICE pointers owned by the caller (row-major, batch-major,
 * contiguous), run asynchronously on the given cudaStream_t (passed as void*)
 * and allocate nothing.  Nothing throws; every function returns
 *   0   success,
 *  <0   bad argument / unsupported shape (QPMPC_B200_E*),
 *  >0   a cudaError_t.
 * Per-instance outcome is written to status[batch]:
 *   0 solved, 1 iteration limit, 2 infeasible, 3 numerical failure (Hessian not
 *   positive definite, or non-finite operands);
 * Plan.is_empty  <=>  status != 0  (qpmpc/plan.py:36,45-48).  U of an
 * unsolved instance is filled with NaN.
 *
 * Operand modes say how each of A, B, C, D, e is laid out:
 *   ABSENT            pointer ignored (C or D "None": the null matrix,
 *                     qpmpc/mpc_problem.py:55-60)
 *   SHARED_LTI        [r, c]          one matrix for all instances and steps
 *   SHARED_LTV        [N, r, c]       per step, shared by all instances
 *   BATCH_LTI         [batch, r, c]   per instance, time-invariant
 *   BATCH_LTV         [batch, N, r, c]
 * with (r, c) = (nx,nx) A, (nx,nu) B, (nc,nx) C, (nc,nu) D, (nc,1) e.
 * x0 is [batch, nx] (BATCH) or [nx] (SHARED); goal likewise; targets is
 * [batch, N*nx] or [N*nx]; goal / targets may be ABSENT.
 */
#ifndef QPMPC_B200_H
#define QPMPC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QPMPC_B200_VERSION 100 /* 0.1.0 */

enum {
    QPMPC_B200_ABSENT = 0,
    QPMPC_B200_SHARED_LTI = 1,
    QPMPC_B200_SHARED_LTV = 2,
    QPMPC_B200_BATCH_LTI = 3,
    QPMPC_B200_BATCH_LTV = 4
};
enum { QPMPC_B200_VEC_ABSENT = 0, QPMPC_B200_VEC_SHARED = 1, QPMPC_B200_VEC_BATCH = 2 };
enum { QPMPC_B200_F64 = 0, QPMPC_B200_F32 = 1 };
/* Solver behind the condensed QP (what qpsolvers.solve_problem's `solver`
 * string selects in the reference, qpmpc/solve_mpc.py:43):
 *   ACTIVE_SET  Goldfarb-Idnani dual active set: exact, every shape.
 *   PDIP        Mehrotra predictor-corrector interior point + active-set
 *               polish (exact when the polish is accepted), stopped at
 *               desc.tol: QPMPC_B200_F64, N*nu <= 32 and N*nc <= 128 only, not
 *               available through qpmpc_b200_solve_scatter
 *               (QPMPC_B200_EUNSUPPORTED otherwise).  It has no infeasibility
 *               certificate: an infeasible instance ends as STATUS_MAX_ITER
 *               or STATUS_NOT_SPD, never as STATUS_SOLVED. */
enum { QPMPC_B200_ACTIVE_SET = 0, QPMPC_B200_PDIP = 1 };
enum { QPMPC_B200_FLAG_NO_POLISH = 1 /* PDIP: skip the polish */ };

enum {
    QPMPC_B200_STATUS_SOLVED = 0,
    QPMPC_B200_STATUS_MAX_ITER = 1,
    QPMPC_B200_STATUS_INFEASIBLE = 2,
    QPMPC_B200_STATUS_NOT_SPD = 3
};

enum {
    QPMPC_B200_EINVAL = -1,      /* null/contradictory argument */
    QPMPC_B200_ESHAPE = -2,      /* N*nu or N*nc beyond the compiled kernels */
    QPMPC_B200_EWEIGHT = -3,     /* w_u <= 0, or neither w_t nor w_x set */
    QPMPC_B200_ENODEVICE = -4,   /* no CUDA device / not sm_100 */
    QPMPC_B200_EUNSUPPORTED = -5 /* combination not implemented */
};

/* Problem descriptor.  Mirrors the constructor of MPCProblem
 * (qpmpc/mpc_problem.py:88-102) plus solver settings. */
typedef struct qpmpc_b200_desc {
    int32_t batch;      /* number of independent MPC instances */
    int32_t N;          /* nb_timesteps */
    int32_t nx, nu, nc; /* state, input, per-step inequality dimensions */
    int32_t dtype;      /* QPMPC_B200_F64 / _F32: type of every operand and of U */
    int32_t mode_A, mode_B, mode_C, mode_D, mode_e;
    int32_t mode_x0, mode_goal, mode_targets;
    int32_t has_wt;     /* terminal_cost_weight is not None */
    int32_t has_wx;     /* stage_state_cost_weight is not None */
    double w_t, w_x, w_u;
    int32_t method;     /* QPMPC_B200_ACTIVE_SET (default, exact) or _PDIP */
    int32_t max_iter;   /* <= 0: library default */
    double tol;         /* PDIP stopping tolerance; <= 0: library default */
    int32_t paired;     /* 1: every C_k, D_k, has rows [M; -M] (two-sided
                           bounds); lets the kernel keep one row per pair.
                           0 is always valid. */
    int32_t flags;      /* QPMPC_B200_FLAG_*; 0 is always valid */
} qpmpc_b200_desc;

/* Inputs of the path (device pointers for the device entry points, host
 * pointers for qpmpc_b200_solve_host).  Element type is desc->dtype. */
typedef struct qpmpc_b200_operands {
    const void *A, *B, *C, *D, *e;
    const void *x0, *goal, *targets;
} qpmpc_b200_operands;

/* Outputs.  U [batch, N*nu] and status [batch] are required; iters [batch]
 * (solver iterations) and Z [batch, N*nc] (multipliers of G u <= h, >= 0) are
 * optional (NULL to skip). */
typedef struct qpmpc_b200_outputs {
    void *U;
    int32_t *status;
    int32_t *iters;
    void *Z;
} qpmpc_b200_outputs;

/* Condense + solve on the device.  Replaces qpmpc/solve_mpc.py:42-43 for a
 * whole batch.  Asynchronous on `stream`. */
int qpmpc_b200_solve(const qpmpc_b200_desc *desc, const qpmpc_b200_operands *in,
                     const qpmpc_b200_outputs *out, void *stream);

/* Condense + solve with a FUSED GATHER: besides (or instead of: out->U and
 * out->status may be NULL) the local outputs, every instance's U row and status
 * are stored by the kernel's epilogue into rows [row_offset, row_offset + batch)
 * of each of `count` destination buffers.  With peer-mapped destinations (CUDA
 * IPC / symmetric memory of the other GPUs of the box) the stores travel over
 * NVLink while the remaining instances are still being solved: this replaces
 * the all-gather of the stacked U trajectories that follows a batch-sharded
 * solve (one rank per GPU, rank r solves rows [r*batch, (r+1)*batch)).  The
 * caller synchronises the ranks afterwards (any barrier). */
typedef struct qpmpc_b200_peers {
    int32_t count;      /* destinations, 1..8 */
    int32_t reserved;
    int64_t row_offset; /* first destination row of this call's instances */
    void *U[8];         /* [total_rows, N*nu] each, dtype of desc */
    int32_t *status[8]; /* [total_rows] each, or NULL */
} qpmpc_b200_peers;

int qpmpc_b200_solve_scatter(const qpmpc_b200_desc *desc, const qpmpc_b200_operands *in,
                             const qpmpc_b200_outputs *out, const qpmpc_b200_peers *peers,
                             void *stream);

/* Same with HOST buffers; synchronises before returning.  This is the call a
 * host program that has no device memory of its own binds.
 *  - PINNED (page-locked) buffers, the fast path: zero-copy.  Page-locked host
 *    memory is mapped into the device's address space, so ONE launch of the
 *    solve kernel does everything: each CTA's bulk-TMA staging pulls its
 *    operands over PCIe, the epilogue stores the U rows (and status) straight
 *    into the caller's buffers.  Upload, compute and download overlap CTA by
 *    CTA; only operands shared by the batch are copied to the device first.
 *    (QPMPC_B200_HOST_ZEROCOPY=0: staged instead -- chunks of the batch whose
 *    upload, kernel and download overlap on three streams, captured once into
 *    a CUDA graph per (descriptor, pointers) and replayed.)
 *  - pageable buffers: staged through device buffers the library caches per
 *    calling thread, copied straight from / to the caller's pointers.
 * Cached buffers, streams and graphs are freed when the calling thread exits
 * or switches to another device. */
int qpmpc_b200_solve_host(const qpmpc_b200_desc *desc, const qpmpc_b200_operands *in,
                          const qpmpc_b200_outputs *out, int device);

/* Condensed QP fields of MPCQP (qpmpc/mpc_qp.py:28-37,93-117), per instance.
 * Any output may be NULL.  P [batch,n,n], q [batch,n], G [batch,m,n],
 * h [batch,m], Phi [batch,N*nx,nx], Psi [batch,N*nx,n], phi_last [batch,nx,nx],
 * psi_last [batch,nx,n]; n = N*nu, m = N*nc. */
typedef struct qpmpc_b200_qp_fields {
    void *P, *q, *G, *h, *Phi, *Psi, *phi_last, *psi_last;
} qpmpc_b200_qp_fields;

int qpmpc_b200_condense(const qpmpc_b200_desc *desc, const qpmpc_b200_operands *in,
                        const qpmpc_b200_qp_fields *out, void *stream);

/* SHARED-MODEL FAST PATH: factor once per model, then per solve only the vectors
 * change.  The batched form of keeping one MPCQP and calling update_cost_vector /
 * update_constraint_vector between solves (qpmpc/mpc_qp.py:129-163): when A, B, C, D
 * are shared by the batch (modes SHARED_LTI / SHARED_LTV, or ABSENT for C, D), the
 * condensed P and G, the Cholesky factor of P, M = G L^-T and the linear maps
 * x0, goal, targets -> q and x0 -> h are the same for every instance and every call.
 *   qpmpc_b200_factor          condenses the model (batch is ignored; in->x0 only has to
 *                              be valid) and writes the record: `record` is a device
 *                              buffer of qpmpc_b200_factor_bytes(desc) bytes.
 *   qpmpc_b200_solve_factored  solves desc->batch instances that share the record: only
 *                              e, x0, goal, targets are read from `in` (same modes,
 *                              weights and dimensions as at factor time).
 * Needs desc->paired rows with nc even, N*nu <= 32 and N*nc <= 64, and the default
 * method; QPMPC_B200_EUNSUPPORTED otherwise (qpmpc_b200_factor_bytes returns 0). */
size_t qpmpc_b200_factor_bytes(const qpmpc_b200_desc *desc);
int qpmpc_b200_factor(const qpmpc_b200_desc *desc, const qpmpc_b200_operands *in, void *record,
                      void *stream);
int qpmpc_b200_solve_factored(const qpmpc_b200_desc *desc, const qpmpc_b200_operands *in,
                              const void *record, const qpmpc_b200_outputs *out, void *stream);

/* X[b, k+1] = A_k X[b, k] + B_k U[b, k], X[b, 0] = x0[b]; X is [batch, N+1, nx].
 * Replaces MPCProblem.integrate (qpmpc/mpc_problem.py:316-335).  Uses
 * desc->mode_A/B/x0, in->A/B/x0; U is [batch, N*nu]. */
int qpmpc_b200_integrate(const qpmpc_b200_desc *desc, const qpmpc_b200_operands *in,
                         const void *U, void *X, void *stream);

/* Receding-horizon closed loop of examples/wheeled_inverted_pendulum.py:99-118
 * for a batch, state resident on the device: per cycle { reference trajectory
 * from the state (get_target_states, examples/...:65-83); condense + solve
 * (qpmpc/solve_mpc.py:42-43); `substeps` plant steps of
 * WheeledInvertedPendulum.integrate (qpmpc/systems/wheeled_inverted_pendulum.py:
 * 127-160) under the first input of the plan }.  The problem must have nx = 4,
 * nu = 1 and per-instance x0 / goal / targets: in->x0 [batch,4] is the state
 * (read and overwritten), in->goal [batch,4] and in->targets [batch,N*4] are
 * work buffers the loop rewrites every cycle; out->U / status / iters hold the
 * plan of the last cycle.  An instance without a plan in some cycle gets input
 * 0 for that cycle and is counted in *unsolved.  With loop->record the whole
 * loop is 2 launches (the instances are independent: each lane group of the
 * shared-model kernel runs all its cycles -- solve, plant, next targets -- without
 * leaving the SM; QPMPC_B200_LOOP_FUSED=0: 2 per cycle); without, 2 * cycles + 1.
 * Asynchronous on `stream`. */
typedef struct qpmpc_b200_closed_loop {
    int32_t cycles;          /* control cycles (200 in BASELINE config 3) */
    int32_t substeps;        /* plant steps per cycle (NB_SUBSTEPS = 15) */
    double dt;               /* plant step = sampling_period / substeps */
    double sampling_period;  /* T of the MPC model */
    double length;           /* pendulum length (omega^2 = gravity / length) */
    double gravity;
    const void *v_target;    /* [batch] target ground velocity, dtype of desc */
    void *trajectory;        /* optional [cycles + 1, batch, 4] states after each cycle */
    int32_t *unsolved;       /* optional device counter, incremented per missing plan */
    const void *record;      /* optional: record of qpmpc_b200_factor for this model -- every
                                cycle then runs qpmpc_b200_solve_factored (the model is
                                condensed and factored once instead of batch x cycles times) */
    int32_t *upright;        /* optional device counter: (instance, cycle) pairs whose pitch
                                was within +-1.2 rad when the cycle's MPC was solved */
    int64_t *iterations;     /* optional [cycles] device array: solver iterations summed over
                                the batch, per cycle (needs out->iters) */
} qpmpc_b200_closed_loop;

int qpmpc_b200_pendulum_closed_loop(const qpmpc_b200_desc *desc, const qpmpc_b200_operands *in,
                                    const qpmpc_b200_outputs *out, const qpmpc_b200_closed_loop *loop,
                                    void *stream);

/* Walking loop of examples/lipm_walking_controller.py:307-335 for a batch, state resident on
 * the device: per cycle { ZMP bounds e_k of the horizon and the goal from the phase machine
 * (update_goal_and_constraints :175-205, PhaseStepper :104-172); condense + solve
 * (qpmpc/solve_mpc.py:42-43); `substeps` constant-jerk integration steps under the first input
 * (:208-226); phase advance and foot switch (:331-334) }.  The problem must have nx = 3,
 * nu = 1, nc = 2 (C = [+zmp; -zmp]) with per-instance x0 [batch,3] (the state, read and
 * overwritten), per-instance per-step e [batch,N,2] (mode_e = BATCH_LTV) and per-instance goal
 * [batch,3]: both are work buffers the loop rewrites every cycle.  An instance without a plan
 * in some cycle gets jerk 0 for that cycle and is counted in *unsolved.  2 * cycles + 1
 * launches, asynchronous on `stream`. */
typedef struct qpmpc_b200_lipm_loop {
    int32_t cycles;          /* control cycles (300 in the reference example) */
    int32_t substeps;        /* integration steps per cycle (15) */
    int32_t nb_dsp_steps;    /* round(dsp_duration / sampling_period) */
    int32_t nb_ssp_steps;    /* round(ssp_duration / sampling_period) */
    double sampling_period;  /* T of the MPC model */
    double foot_size;
    double max_zmp_dist;     /* bound written where the ZMP is unconstrained (MAX_ZMP_DIST) */
    void *support_foot;      /* [batch] position of the support foot, read and updated */
    const void *strides;     /* [batch, 2] alternating strides */
    int32_t *phase_index;    /* [batch] PhaseStepper.index, read and updated */
    int32_t *stride_index;   /* [batch] PhaseStepper.stride_index, read and updated */
    void *trajectory;        /* optional [cycles + 1, batch, 3] states after each cycle */
    int32_t *unsolved;       /* optional device counter, incremented per missing plan */
    const void *record;      /* optional: record of qpmpc_b200_factor for this model */
} qpmpc_b200_lipm_loop;

int qpmpc_b200_lipm_closed_loop(const qpmpc_b200_desc *desc, const qpmpc_b200_operands *in,
                                const qpmpc_b200_outputs *out, const qpmpc_b200_lipm_loop *loop,
                                void *stream);

/* Scratch the device entry points need from the caller: none (0).  Shapes whose
 * matrices exceed shared memory (n > 72 in fp64 with m = 2 n) take a workspace
 * the library allocates and frees in stream order (cudaMallocAsync) itself. */
size_t qpmpc_b200_workspace_bytes(const qpmpc_b200_desc *desc);

/* Largest n = N*nu and m = N*nc the kernels accept for this dtype (n <= 512,
 * m <= 4096; beyond 227 KB of matrices the CTA kernel works out of global memory). */
int qpmpc_b200_max_vars(int dtype);
int qpmpc_b200_max_rows(int dtype, int n);

/* Measures the FP64 FMA peak of `device` (TFLOP/s) with a register-only
 * kernel: the denominator of the FP64 roofline bench.py reports. */
int qpmpc_b200_fp64_peak(int device, double *tflops);

/* Number of kernels this library has launched in this process. */
long long qpmpc_b200_launch_count(void);

const char *qpmpc_b200_strerror(int code);
int qpmpc_b200_version(void);

#ifdef __cplusplus
}
#endif
#endif /* QPMPC_B200_H */
