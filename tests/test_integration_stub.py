"""The reference-side binding printed in INTEGRATION.md section 2 is code, not prose: the block is
extracted from the document and executed -- on any machine up to the point where it binds the
library (struct layouts against the header's ctypes mirror), on a GPU all the way through
``solve(problem)`` against ``solve_mpc``."""

import ctypes
import os
import re

import numpy as np
import pytest

from conftest import golden_problem, load_golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _stub_namespace():
    from qpmpc_b200 import build

    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    blocks = re.findall(r"```python\n(.*?)```", text, flags=re.S)
    stub = next(b for b in blocks if "qpmpc/b200_backend.py" in b)
    stub = stub.replace('ctypes.CDLL("libqpmpc_b200.so")', f'ctypes.CDLL({build.build_library()!r})')
    ns = {}
    exec(compile(stub, "INTEGRATION.md:b200_backend", "exec"), ns)
    return ns


def test_stub_binds_the_library_with_the_headers_layouts():
    from qpmpc_b200 import _capi

    ns = _stub_namespace()
    for name, mirror in (("Desc", _capi.Desc), ("Operands", _capi.Operands), ("Outputs", _capi.Outputs)):
        assert ctypes.sizeof(ns[name]) == ctypes.sizeof(mirror)
        assert [(f[0], ctypes.sizeof(f[1])) for f in ns[name]._fields_] == \
               [(f[0], ctypes.sizeof(f[1])) for f in mirror._fields_]
    assert ns["_lib"].qpmpc_b200_solve_host.argtypes is not None


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["triple_integrator", "humanoid", "pendulum", "random_ltv_cd"])
def test_stub_solves_like_solve_mpc(name):
    from qpmpc_b200 import solve_mpc

    ns = _stub_namespace()
    problem = golden_problem(load_golden(name))
    sol = ns["solve"](problem)
    plan = solve_mpc(problem, "b200")
    assert sol.found and not plan.is_empty
    assert np.abs(sol.x - plan.qpsol.x).max() <= 1e-9 * max(1.0, np.abs(plan.qpsol.x).max())
    infeasible = golden_problem(load_golden("triple_integrator"))
    infeasible.update_initial_state(np.array([0.0, 0.0, 5.0]))
    assert not ns["solve"](infeasible).found
