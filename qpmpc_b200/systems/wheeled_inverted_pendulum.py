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
        a_max = self.max_ground_accel
        return MPCProblem(
            transition_state_matrix=A,
            transition_input_matrix=B,
            ineq_state_matrix=None,
            ineq_input_matrix=np.array([[1.0], [-1.0]]),
            ineq_vector=np.array([a_max, a_max], dtype=float),
            nb_timesteps=self.nb_timesteps,
            terminal_cost_weight=terminal_cost_weight,
            stage_state_cost_weight=stage_state_cost_weight,
            stage_input_cost_weight=stage_input_cost_weight,
        )

    def integrate(self, state: np.ndarray, ground_accel, dt: float) -> np.ndarray:
        """One plant step of duration dt under a constant ground acceleration."""
        r, theta, rd, thetad = state
        rdd = ground_accel
        thetadd = self.omega**2 * (
            np.sin(theta) - (rdd / self.GRAVITY) * np.cos(theta)
        )
        r_next = r + dt * (rd + dt * (rdd / 2))
        rd_next = rd + dt * rdd
        theta_next = theta + dt * (thetad + dt * (thetadd / 2))
        thetad_next = thetad + dt * thetadd
        return np.array([r_next, theta_next, rd_next, thetad_next]).flatten()
