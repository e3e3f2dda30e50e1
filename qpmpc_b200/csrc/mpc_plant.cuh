// mpc_plant.cuh -- the pieces of the receding-horizon loop of
// examples/wheeled_inverted_pendulum.py:99-118 (reference tree) that sit
// between two MPC solves, batched, state resident on the GPU:
//   * the plant: NB_SUBSTEPS calls of WheeledInvertedPendulum.integrate
//     (qpmpc/systems/wheeled_inverted_pendulum.py:127-160) under the first
//     input of the plan (examples/...:110-111);
//   * the reference trajectory of the next cycle: get_target_states
//     (examples/...:65-83) -> goal state (last nx entries) and stage targets.
// One thread per instance; state is [batch, 4] = [r, theta, r_dot, theta_dot].
#pragma once

#include "mpc_common.cuh"

namespace qpmpc {

// ---- the plant pieces as device functions: the step kernels below and the fused loop of the
// shared-model solve kernel (mpc_kernels.cuh, SolveParams::loop) run the same code ------------

// NB substeps of the second-order Taylor step of the nonlinear dynamics (systems/...:150-160).
// The pitch moves by a small angle per substep, so sin / cos of the new pitch come from the old
// ones by the angle-addition formulas with sin / cos of the INCREMENT from their Taylor
// polynomials (|increment| <= 0.05: truncation below 1e-17, i.e. double rounding); one full
// sincos per call -- and per large increment, a pendulum that has fallen -- instead of one per
// substep.  (Profiled: sincos was 45 % of the instructions of the fused config-3 loop.)
template <typename T>
__device__ __forceinline__ void pendulum_integrate(T &r, T &th, T &rd, T &thd, T u, int substeps, T dt, T w2, T g) {
    T sn, cs;
    sincos_(th, &sn, &cs);
    const T rdd = u;
    const T ug = rdd / g, hu = rdd / T(2);  // (the input is constant over the cycle: one division)
    for (int s = 0; s < substeps; ++s) {
        const T thdd = w2 * (sn - ug * cs);
        r = r + dt * (rd + dt * hu);
        rd = rd + dt * rdd;
        const T dth = dt * (thd + dt * (thdd * T(0.5)));
        th = th + dth;
        thd = thd + dt * thdd;
        if (s + 1 < substeps) {
            if (abs_(dth) <= T(0.05)) {
                const T d2 = dth * dth;
                const T sd = dth * (T(1) + d2 * (T(-1.0 / 6.0) + d2 * (T(1.0 / 120.0) + d2 * T(-1.0 / 5040.0))));
                const T cd = T(1) + d2 * (T(-0.5) + d2 * (T(1.0 / 24.0) + d2 * (T(-1.0 / 720.0) + d2 * T(1.0 / 40320.0))));
                const T ns = sn * cd + cs * sd;
                cs = cs * cd - sn * sd;
                sn = ns;
            } else {
                sincos_(th, &sn, &cs);
            }
        }
    }
}
// get_target_states: position ramp r + k T v, velocity v, zero pitch (examples/...:77-82)
template <typename T>
__device__ __forceinline__ void pendulum_target(T *tg, T r, T v, int k, T Ts) {
    tg[0] = r + ((T)k * Ts) * v;
    tg[1] = T(0);
    tg[2] = v;
    tg[3] = T(0);
}
// constant-jerk integration (examples/lipm_walking_controller.py:219-225)
template <typename T>
__device__ __forceinline__ void lipm_integrate(T &pos, T &vel, T &acc, T jerk, int substeps, T dt) {
    const T j6 = dt * jerk / T(6), j2 = jerk / T(2);  // (constant over the cycle: one division)
    for (int s = 0; s < substeps; ++s) {
        const T p1 = pos + dt * (vel + dt * (acc * T(0.5) + j6));
        const T v1 = vel + dt * (acc + dt * j2);
        acc = acc + dt * jerk;
        pos = p1;
        vel = v1;
    }
}
// get_nb_steps (:132-163): lengths of the first five phases that cover the horizon (the sixth
// takes what is left: at most nb_ssp steps for horizons of two steps)
struct LipmPhases {
    int n0, n1, n2, n3, n4;
};
__device__ __forceinline__ LipmPhases lipm_phases(int index, int nb_dsp, int nb_ssp, int N) {
    LipmPhases q;
    int off = index;
    q.n0 = max(0, nb_dsp - off);
    off = max(0, off - nb_dsp);
    q.n1 = max(0, nb_ssp - off);
    int rem = N - q.n0 - q.n1;
    q.n2 = min(nb_dsp, rem);
    rem = max(0, rem - nb_dsp);
    q.n3 = min(nb_ssp, rem);
    rem = max(0, rem - nb_ssp);
    q.n4 = min(nb_dsp, rem);
    return q;
}
// update_goal_and_constraints (:175-205): ZMP bounds of step k (hi: z <= hi, lo: -z <= lo)
template <typename T>
__device__ __forceinline__ void lipm_bounds(const LipmPhases &q, int k, T foot, T next, T last, T hf, T big, T &hi,
                                            T &lo) {
    hi = big, lo = big;
    int j = k;
    if (j < q.n0) {
    } else if ((j -= q.n0) < q.n1) {
        hi = foot + hf, lo = -(foot - hf);
    } else if ((j -= q.n1) < q.n2) {
    } else if ((j -= q.n2) < q.n3) {
        hi = next + hf, lo = -(next - hf);
    } else if ((j -= q.n3) < q.n4) {
    } else {
        hi = last + hf, lo = -(last - hf);
    }
}

struct PendulumStepParams {
    int batch, N, n;      // instances, horizon, N * nu (row stride of U)
    int substeps;         // plant steps per control cycle (0: only write targets)
    double dt, T;         // plant step and MPC sampling period
    double omega2, g;     // g / length, g
    void *state;          // [batch, 4] in/out
    const void *U;        // [batch, n] plan of the cycle that just ended
    const int *status;    // [batch] 0 = solved; otherwise the input is 0
    const void *v_target; // [batch] target ground velocity
    void *goal;           // [batch, 4] out
    void *targets;        // [batch, N * 4] out
    void *traj;           // optional [batch, 4] slot of the recorded trajectory
    int *unsolved;        // optional counter of (instance, cycle) pairs without a plan
    int *upright;         // optional counter of (instance, cycle) pairs solved with |pitch| <= 1.2 rad
    const int *iters;     // optional [batch] solver iterations of the cycle that just ended ...
    long long *iter_sum;  // ... summed over the batch into this slot
};

template <typename T>
__global__ void __launch_bounds__(128) pendulum_step_kernel(const PendulumStepParams p) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= p.batch) return;
    T *st = static_cast<T *>(p.state) + (size_t)b * 4;
    T r = st[0], th = st[1], rd = st[2], thd = st[3];
    if (p.substeps > 0) {
        const bool ok = p.status[b] == 0;
        const T u = ok ? static_cast<const T *>(p.U)[(size_t)b * p.n] : T(0);
        if (!ok && p.unsolved) atomicAdd(p.unsolved, 1);
        // the MPC of this cycle was solved from the state before the plant moves
        if (p.upright && abs_(th) <= T(1.2)) atomicAdd(p.upright, 1);
        if (p.iter_sum && p.iters) atomicAdd(reinterpret_cast<unsigned long long *>(p.iter_sum),
                                             (unsigned long long)p.iters[b]);
        pendulum_integrate<T>(r, th, rd, thd, u, p.substeps, (T)p.dt, (T)p.omega2, (T)p.g);
        st[0] = r, st[1] = th, st[2] = rd, st[3] = thd;
    }
    if (p.traj) {
        T *tr = static_cast<T *>(p.traj) + (size_t)b * 4;
        tr[0] = r, tr[1] = th, tr[2] = rd, tr[3] = thd;
    }
    // get_target_states: position ramp r + k T v, velocity v, zero pitch (examples/...:77-82)
    const T v = static_cast<const T *>(p.v_target)[b];
    T *tg = static_cast<T *>(p.targets) + (size_t)b * p.N * 4;
    for (int k = 0; k < p.N; ++k) pendulum_target<T>(tg + k * 4, r, v, k, (T)p.T);
    pendulum_target<T>(static_cast<T *>(p.goal) + (size_t)b * 4, r, v, p.N, (T)p.T);
}

// ---------------------------------------------------------------------------
// The walking loop of examples/lipm_walking_controller.py:307-335 (reference tree), batched:
// between two MPC solves the state is integrated under the first jerk of the plan
// (integrate, :208-226), the phase machine advances (PhaseStepper.advance, :124-130, and the
// foot switch of :331-334), and the next cycle's LTV constraint vector e_k and goal are written
// (update_goal_and_constraints, :175-205, with PhaseStepper.get_nb_steps, :132-163).
// One thread per instance; state is [batch, 3] = [pos, vel, accel].
// ---------------------------------------------------------------------------
struct LipmStepParams {
    int batch, N, n;
    int substeps;          // integration substeps per control cycle (0: only write e and the goal)
    int nb_dsp, nb_ssp;    // steps of a double / single support phase
    double dt;             // integration step = sampling period / substeps
    double foot_size, max_zmp;
    void *state;           // [batch, 3] in/out
    const void *U;         // [batch, n] plan of the cycle that just ended
    const int *status;     // [batch] 0 = solved; otherwise the jerk is 0
    void *support_foot;    // [batch] in/out: position of the support foot
    const void *strides;   // [batch, 2]
    int *phase_index;      // [batch] in/out
    int *stride_index;     // [batch] in/out
    void *e;               // [batch, N, 2] out: ZMP bounds of the next cycle
    void *goal;            // [batch, 3] out
    void *traj;            // optional [batch, 3] slot of the recorded trajectory
    int *unsolved;
};

template <typename T>
__global__ void __launch_bounds__(128) lipm_step_kernel(const LipmStepParams p) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= p.batch) return;
    T *st = static_cast<T *>(p.state) + (size_t)b * 3;
    T pos = st[0], vel = st[1], acc = st[2];
    int index = p.phase_index[b], sidx = p.stride_index[b];
    T foot = static_cast<T *>(p.support_foot)[b];
    const T *strides = static_cast<const T *>(p.strides) + (size_t)b * 2;
    const int cyc = p.nb_dsp + p.nb_ssp;
    if (p.substeps > 0) {
        const bool ok = p.status[b] == 0;
        const T jerk = ok ? static_cast<const T *>(p.U)[(size_t)b * p.n] : T(0);
        if (!ok && p.unsolved) atomicAdd(p.unsolved, 1);
        lipm_integrate<T>(pos, vel, acc, jerk, p.substeps, (T)p.dt);
        st[0] = pos, st[1] = vel, st[2] = acc;
        // PhaseStepper.advance and the foot switch (:331-334)
        index = index + 1 >= cyc ? 0 : index + 1;
        if (index == 0) {
            foot = foot + strides[sidx];
            sidx = (sidx + 1) % 2;
            static_cast<T *>(p.support_foot)[b] = foot;
            p.stride_index[b] = sidx;
        }
        p.phase_index[b] = index;
    }
    if (p.traj) {
        T *tr = static_cast<T *>(p.traj) + (size_t)b * 3;
        tr[0] = pos, tr[1] = vel, tr[2] = acc;
    }
    const LipmPhases ph = lipm_phases(index, p.nb_dsp, p.nb_ssp, p.N);
    const T next = foot + strides[sidx];
    const T last = next + strides[(sidx + 1) % 2];
    const T hf = T(0.5) * (T)p.foot_size, big = (T)p.max_zmp;
    T *e = static_cast<T *>(p.e) + (size_t)b * p.N * 2;
    for (int k = 0; k < p.N; ++k) lipm_bounds<T>(ph, k, foot, next, last, hf, big, e[2 * k], e[2 * k + 1]);
    T *gl = static_cast<T *>(p.goal) + (size_t)b * 3;
    gl[0] = ph.n4 > 0 ? last : next;
    gl[1] = T(0);
    gl[2] = T(0);
}

}  // namespace qpmpc
