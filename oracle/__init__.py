"""CPU oracle for the condense + QP-solve path.  TEST INFRASTRUCTURE ONLY.

Nothing under ``qpmpc_b200/`` imports this package.  It is used by ``tests/``,
by ``__graft_entry__.smoke()`` and by ``bench.py``'s CPU-baseline / reference
arm -- as the checker and the timed CPU leg, never as the product.

Pieces:
  * :mod:`oracle.condense_np` -- NumPy restatement of the reference condensing
    (``qpmpc/mpc_qp.py:39-163``), PINNED field by field against the reference's
    own ``MPCQP`` through ``tests/golden/``.
  * ``mpc_oracle.c`` (loaded here through ctypes) -- the same condensing in C,
    a Goldfarb-Idnani dual active-set QP solver (quadprog's published
    algorithm) and a KKT certifier, batched with OpenMP.
  * :mod:`oracle.pdip_np` -- NumPy model of the interior-point + polish
    iteration the CUDA kernel runs, used to cross-check the active-set solver.

Solver half: pinned to PUBLISHED optima (quadprog's documented example, the
qpsolvers README example, Hock-Schittkowski 21/35/76/118/268 of the
Maros-Meszaros set: ``tests/published_qps.py``), not to outputs of the proxqp /
quadprog wheels themselves (not installable offline; the reference tests hold
no numerical QP answer beyond ``U = 0``); see the header of ``mpc_oracle.c``.
"""

from .capi import (  # noqa: F401
    STATUS_OK,
    build,
    condense,
    kkt,
    num_threads,
    qp_gi,
    solve_batch,
)
