"""The reference's OWN example scripts, run unmodified against this package.

``/root/reference/examples/*.py`` are the closed-loop programs a user of the reference has:
they import ``qpmpc``, build an ``MPCProblem`` and call ``solve_mpc(problem, solver="proxqp")``
once (triple integrator, humanoid step) or once per control cycle (wheeled inverted pendulum:
100 cycles; LIPM walking controller: 300 cycles with the constraint vector rewritten every
cycle).  Here each script is executed as ``__main__`` straight from the reference tree -- nothing
is copied -- with the repo's ``qpmpc`` alias on the path and headless stand-ins for what is out
of scope (plotting: ``pylab``, ``qpmpc.live_plots``; pacing: ``loop_rate_limiters``; the
``qpsolvers`` package when it is not installed; ``input()``).  The last plan and the states the
script handed to its live plot are checked.

On a CPU-only machine the C-ABI entry points are served by the device source compiled for the
host (``emulated_engine``); with a GPU (and the reference tree present) the CUDA library runs.
"""

import builtins
import os
import runpy
import sys
import types

import numpy as np
import pytest

REF_EXAMPLES = os.environ.get("QPMPC_REFERENCE_EXAMPLES", "/root/reference/examples")
pytestmark = pytest.mark.skipif(not os.path.isdir(REF_EXAMPLES), reason="the reference tree is not present on this machine")


class _Anything:
    """Accepts every attribute access, call and item assignment (a headless figure)."""

    def __init__(self, *a, **k):
        pass

    def __getattr__(self, name):
        return _Anything()

    def __call__(self, *a, **k):
        return _Anything()

    def __iter__(self):
        return iter(())


class _Recorder(_Anything):
    """Stands in for the live plots: remembers what the script shows."""

    log = []

    def update(self, *a, **k):
        if "state" in k:
            _Recorder.log.append(np.array(k["state"], dtype=float))
        return None

    def update_line(self, name, xs, ys):
        if name == "cur_pos":
            _Recorder.log.append(np.array([ys[0]], dtype=float))


def _headless(monkeypatch):
    def module(name, **attrs):
        mod = types.ModuleType(name)
        mod.__dict__.update(attrs)
        mod.__getattr__ = lambda attr: _Anything()
        monkeypatch.setitem(sys.modules, name, mod)
        return mod

    module("pylab")
    module("loop_rate_limiters", RateLimiter=_Anything)
    try:
        import qpsolvers  # noqa: F401
    except ImportError:
        module("qpsolvers", available_solvers=["proxqp", "quadprog", "osqp"])
    import qpmpc  # the alias package at the repo root (opts in to serving solver="proxqp")

    assert qpmpc.MPCProblem.__module__.startswith("qpmpc_b200")
    plots = module("qpmpc.live_plots", WheeledInvertedPendulumPlot=_Recorder, LivePlot=_Recorder)
    module("qpmpc.live_plots.live_plot", LivePlot=_Recorder)
    monkeypatch.setattr(qpmpc, "live_plots", plots, raising=False)
    monkeypatch.setattr(builtins, "input", lambda *a: "")
    monkeypatch.setattr(sys, "argv", ["example"])
    _Recorder.log = []


def _run(name, monkeypatch):
    _headless(monkeypatch)
    return runpy.run_path(os.path.join(REF_EXAMPLES, name), run_name="__main__")


def _check(name, g):
    plan = g["plan"]
    assert not plan.is_empty
    if name == "triple_integrator.py":
        X = plan.states
        # bang-bang plan under the acceleration limit; end position of the known optimum (SURVEY's U*)
        assert abs(X[-1, 0] - 0.71959) < 1e-4 and np.abs(X[:, 2]).max() <= 3.0 + 1e-9
    elif name == "humanoid_one_step.py":
        assert plan.states.shape[1] == 3 and np.isfinite(plan.states).all()
    elif name == "wheeled_inverted_pendulum.py":
        states = np.stack(_Recorder.log)
        assert len(states) == 100 * 15                                              # 100 control cycles x 15 substeps
        assert np.abs(states[:, 1]).max() < 0.5 and abs(states[-1, 2] - 0.5) < 0.05  # upright, at the target velocity
    else:
        pos = np.concatenate(_Recorder.log)
        assert len(pos) == 300 * 15 and np.isfinite(pos).all()
        assert np.abs(pos).max() < 0.3 and np.abs(pos[-15 * 8:]).max() > 0.005       # sways between the feet, keeps walking


EXAMPLES = ("triple_integrator.py", "humanoid_one_step.py", "wheeled_inverted_pendulum.py", "lipm_walking_controller.py")


@pytest.mark.parametrize("name", EXAMPLES)
def test_reference_example_on_the_host_emulator(name, emulated_engine, monkeypatch):
    g = _run(name, monkeypatch)
    _check(name, g)
    assert emulated_engine.calls >= (1 if "one_step" in name or "triple" in name else 100)


@pytest.mark.gpu
@pytest.mark.parametrize("name", EXAMPLES)
def test_reference_example_on_the_device(name, monkeypatch):
    from qpmpc_b200 import _capi

    before = _capi.launch_count()
    _check(name, _run(name, monkeypatch))
    assert _capi.launch_count() > before
