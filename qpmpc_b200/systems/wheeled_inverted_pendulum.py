"""Wheeled inverted pendulum: linearised MPC model + nonlinear plant step.

Behavioural mirror of ``qpmpc/systems/wheeled_inverted_pendulum.py:16-160``.
State x = [r, theta, r_dot, theta_dot]; input u = ground acceleration.  The
MPC model is the exact zero-order-hold discretisation of the linearised
dynamics  theta_ddot = omega^2 (theta - u / g)  over one sampling period; the
plant step is a second-order Taylor expansion of the nonlinear dynamics.
"""

from typing import Optional

import numpy as np

from ..mpc_problem import MPCProblem


class WheeledInvertedPendulum:
    """Cart-pole-like model of a wheeled biped such as Upkie."""

    GRAVITY: float = 9.81  # m/s^2
    INPUT_DIM: int = 1
    STATE_DIM: int = 4

    def __init__(
        self,
        length: float = 0.6,
        max_ground_accel: float = 10.0,
        nb_timesteps: int = 12,
        sampling_period: float = 0.1,
    ):
        self.length = length
        self.max_ground_accel = max_ground_accel
        self.nb_timesteps = nb_timesteps
        self.sampling_period = sampling_period

    @property
    def omega(self) -> float:
        """Natural frequency sqrt(g / l) of the pendulum."""
        return np.sqrt(self.GRAVITY / self.length)

    @property
    def horizon_duration(self) -> float:
        """Preview window length in seconds."""
        return self.sampling_period * self.nb_timesteps

    def discretized_dynamics(self):
        """Return (A, B) of the ZOH-discretised linear model."""
        T, w, g = self.sampling_period, self.omega, self.GRAVITY
        ch, sh = np.cosh(T * w), np.sinh(T * w)
        A = np.array(
            [
                [1.0, 0.0, T, 0.0],
                [0.0, ch, 0.0, sh / w],
                [0.0, 0.0, 1.0, 0.0],
                [0.0, w * sh, 0.0, ch],
            ]
        )
        B = np.array([[T**2 / 2.0], [-ch / g + 1.0 / g], [T], [-w * sh / g]])
        return A, B

    def build_mpc_problem(
        self,
        stage_input_cost_weight: float = 1e-3,
        stage_state_cost_weight: Optional[float] = None,
        terminal_cost_weight: Optional[float] = 1.0,
    ) -> MPCProblem:
        """MPC problem with |u| <= max_ground_accel and no state constraint."""
        A, B = self.discretized_dynamics()
        bound = float(self.max_ground_accel)
        # two-sided input bound  -bound <= u <= bound  as  [1; -1] u <= [bound; bound]
        D, e = np.array([[1.0], [-1.0]]), np.array([bound, bound])
        return MPCProblem(A, B, None, D, e, self.nb_timesteps, terminal_cost_weight,
                          stage_state_cost_weight, stage_input_cost_weight)

    def integrate(self, state: np.ndarray, ground_accel, dt: float) -> np.ndarray:
        """One plant step of duration dt under a constant ground acceleration."""
        pos, pitch, vel, pitch_rate = np.asarray(state, dtype=float).flatten()
        accel = float(np.asarray(ground_accel).reshape(-1)[0])
        # theta_ddot = omega^2 (sin(theta) - (u / g) cos(theta)); second-order Taylor step
        pitch_accel = self.omega**2 * (np.sin(pitch) - (accel / self.GRAVITY) * np.cos(pitch))
        half = 0.5 * dt * dt
        return np.array([
            pos + dt * vel + half * accel,
            pitch + dt * pitch_rate + half * pitch_accel,
            vel + dt * accel,
            pitch_rate + dt * pitch_accel,
        ])
