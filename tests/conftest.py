"""pytest configuration: the ``gpu`` marker and shared helpers.

``-m "not gpu"`` runs here (no GPU): oracle vs golden vectors, host logic,
C-ABI exports, gloo sharding.  ``-m gpu`` runs on a B200: parity of the CUDA
path (through the C ABI) against the oracle and the golden fixtures.
Nothing here reads /root/reference at run time.
"""

import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch

        have_gpu = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        have_gpu = False
    if have_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    """One tests/golden/condense_<name>.npz as a dict with None for absent fields."""
    d = dict(np.load(os.path.join(GOLDEN, f"condense_{name}.npz")))
    out = {}
    for k, v in d.items():
        out[k] = v
    for k in ("C", "D", "goal", "targets"):
        if out[k].size == 0:
            out[k] = None
    out["N"] = int(out["N"])
    out["w_t"] = None if np.isnan(out["w_t"]) else float(out["w_t"])
    out["w_x"] = None if np.isnan(out["w_x"]) else float(out["w_x"])
    out["w_u"] = float(out["w_u"])
    return out


GOLDEN_NAMES = [
    "triple_integrator", "triple_integrator_stage", "triple_integrator_tiny_wt",
    "humanoid", "pendulum", "random_ltv_cd", "random_ltv_c", "random_ltv_d",
    "triple_integrator_N8", "triple_integrator_N32", "triple_integrator_N64",
]


def golden_problem(g):
    """Rebuild the host-side MPCProblem of a golden fixture (LTV -> lists)."""
    from qpmpc_b200 import MPCProblem

    def op(name):
        arr = g[name]
        if arr is None:
            return None
        return [a for a in arr] if bool(g[f"{name}_ltv"]) else arr

    prob = MPCProblem(
        transition_state_matrix=op("A"), transition_input_matrix=op("B"),
        ineq_state_matrix=op("C"), ineq_input_matrix=op("D"), ineq_vector=op("e"),
        nb_timesteps=g["N"], terminal_cost_weight=g["w_t"],
        stage_state_cost_weight=g["w_x"], stage_input_cost_weight=g["w_u"],
        initial_state=g["x0"], goal_state=g["goal"],
    )
    if g["targets"] is not None:
        prob.update_target_states(g["targets"])
    return prob


@pytest.fixture
def emulated_engine(monkeypatch):
    """The host-side surface (MPCProblem / MPCQP / solve_mpc / Plan) on a machine without a GPU:
    the C-ABI entry points it calls are served by the DEVICE SOURCE compiled for the host
    (tests/emu) on CPU tensors.  Only tests use this seam; the product has no CPU path."""
    import contextlib
    import ctypes

    import torch

    import emu
    from qpmpc_b200 import _capi, batched

    fake = emu.EmulatedLibrary()
    monkeypatch.setattr(_capi, "load", lambda: fake)
    monkeypatch.setattr(batched, "_require_cuda", lambda device: torch.device("cpu"))
    monkeypatch.setattr(batched, "_device_guard", lambda device: contextlib.nullcontext())
    monkeypatch.setattr(batched, "_stream_ptr", lambda device: ctypes.c_void_p(0))
    return fake
