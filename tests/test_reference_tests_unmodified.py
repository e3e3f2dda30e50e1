"""The reference's OWN test files, run unmodified against this package.

``/root/reference/tests/*.py`` import ``qpmpc`` (``from qpmpc import MPCProblem, solve_mpc``,
``from qpmpc.solve_mpc import MPCQP``, ``from qpmpc.systems import WheeledInvertedPendulum``) and
call ``solve_mpc(problem, solver="proxqp")``.  Here they are loaded from the reference tree as they
are -- nothing is copied -- with the repo's ``qpmpc`` compatibility alias on the path, and run

  * on a B200 through the CUDA library (``-m gpu``), and
  * on a CPU-only machine with the C-ABI entry points served by the device source compiled for
    the host (the ``emulated_engine`` fixture): the same kernels, thread by thread.

The reference tree exists in the development container only (it is not shipped to the GPU box):
where it is absent these tests are skipped, and ``tests/test_reference_suite.py`` holds the same
five cases restated.
"""

import importlib.util
import os
import sys
import unittest

import pytest

REF_TESTS = os.environ.get("QPMPC_REFERENCE_TESTS", "/root/reference/tests")
FILES = ("test_humanoid_one_step.py", "test_update_constraint_vector.py", "test_wheeled_inverted_pendulum.py")

pytestmark = pytest.mark.skipif(not all(os.path.exists(os.path.join(REF_TESTS, f)) for f in FILES),
                                reason="the reference tree is not present on this machine")


def _run_unmodified(name):
    import qpmpc  # the alias package at the repo root (opts in to serving solver="proxqp")

    assert qpmpc.MPCProblem.__module__.startswith("qpmpc_b200")  # not the reference's own package
    spec = importlib.util.spec_from_file_location("reference_" + name[:-3], os.path.join(REF_TESTS, name))
    module = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(module)
    suite = unittest.defaultTestLoader.loadTestsFromModule(module)
    assert suite.countTestCases() >= 1
    result = unittest.TextTestRunner(stream=sys.stderr, verbosity=0).run(suite)
    assert result.wasSuccessful(), (result.failures, result.errors)
    return result.testsRun


@pytest.mark.parametrize("name", FILES)
def test_reference_test_file_on_the_host_emulator(name, emulated_engine):
    ran = _run_unmodified(name)
    assert ran >= 1 and emulated_engine.calls >= 1  # the kernels (emulated) did the work


@pytest.mark.gpu
@pytest.mark.parametrize("name", FILES)
def test_reference_test_file_on_the_device(name):
    from qpmpc_b200 import _capi

    before = _capi.launch_count()
    assert _run_unmodified(name) >= 1
    assert _capi.launch_count() > before
