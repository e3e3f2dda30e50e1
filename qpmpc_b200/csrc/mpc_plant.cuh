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
        const T dt = (T)p.dt, w2 = (T)p.omega2, g = (T)p.g;
        for (int s = 0; s < p.substeps; ++s) {
            // second-order Taylor step of the nonlinear dynamics (systems/...:150-160)
            const T rdd = u;
            const T thdd = w2 * (sin(th) - (rdd / g) * cos(th));
            r = r + dt * (rd + dt * (rdd / T(2)));
            rd = rd + dt * rdd;
            th = th + dt * (thd + dt * (thdd / T(2)));
            thd = thd + dt * thdd;
        }
        st[0] = r, st[1] = th, st[2] = rd, st[3] = thd;
    }
    if (p.traj) {
        T *tr = static_cast<T *>(p.traj) + (size_t)b * 4;
        tr[0] = r, tr[1] = th, tr[2] = rd, tr[3] = thd;
    }
    // get_target_states: position ramp r + k T v, velocity v, zero pitch (examples/...:77-82)
    const T v = static_cast<const T *>(p.v_target)[b];
    T *tg = static_cast<T *>(p.targets) + (size_t)b * p.N * 4;
    for (int k = 0; k < p.N; ++k) {
        tg[k * 4 + 0] = r + ((T)k * (T)p.T) * v;
        tg[k * 4 + 1] = T(0);
        tg[k * 4 + 2] = v;
        tg[k * 4 + 3] = T(0);
    }
    T *gl = static_cast<T *>(p.goal) + (size_t)b * 4;
    gl[0] = r + ((T)p.N * (T)p.T) * v;
    gl[1] = T(0);
    gl[2] = v;
    gl[3] = T(0);
}

}  // namespace qpmpc
