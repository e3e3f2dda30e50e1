"""``qpmpc`` compatibility alias of :mod:`qpmpc_b200`.

Code written against stephane-caron/qpmpc -- ``from qpmpc import MPCProblem,
solve_mpc``, ``from qpmpc.solve_mpc import MPCQP``, ``from qpmpc.systems import
WheeledInvertedPendulum`` (``tests/test_humanoid_one_step.py:11-13``,
``tests/test_wheeled_inverted_pendulum.py:11-12`` of the reference) -- imports
this package instead when the repo root is on ``sys.path`` and gets the B200
engine behind the same names (``qpmpc/__init__.py:9-21`` of the reference).

Importing the alias IS the opt-in that lets qpsolvers backend names
(``solver="proxqp"`` ...) be served by the CUDA engine when ``qpsolvers`` is
not installed (:func:`qpmpc_b200.solve_mpc.serve_qpsolvers_names`): that is
what reference call sites pass.  Nothing is computed here: every name is a
re-export, every submodule an alias of the ``qpmpc_b200`` module of that name.
"""

import importlib
import sys

import qpmpc_b200
from qpmpc_b200 import MPCQP, MPCProblem, Plan, solve_mpc
from qpmpc_b200.solve_mpc import serve_qpsolvers_names

__version__ = "3.1.0"  # the reference version whose surface is mirrored
__all__ = ["MPCProblem", "MPCQP", "Plan", "solve_mpc"]

for _name in ("exceptions", "mpc_problem", "mpc_qp", "plan", "solve_mpc", "systems",
              "systems.wheeled_inverted_pendulum"):
    sys.modules[f"{__name__}.{_name}"] = importlib.import_module(f"qpmpc_b200.{_name}")
exceptions = sys.modules[f"{__name__}.exceptions"]
systems = sys.modules[f"{__name__}.systems"]
mpc_problem = sys.modules[f"{__name__}.mpc_problem"]
mpc_qp = sys.modules[f"{__name__}.mpc_qp"]
plan = sys.modules[f"{__name__}.plan"]
# ``qpmpc.solve_mpc`` stays the function, as in the reference (its __init__ rebinds the name);
# ``from qpmpc.solve_mpc import MPCQP`` resolves through sys.modules.

serve_qpsolvers_names(True)
