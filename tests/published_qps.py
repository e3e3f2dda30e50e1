"""Small strictly convex QPs whose optimum is PUBLISHED, in the form the path's
solver takes (``qpsolvers.solve_problem(P, q, G, h)``, ``qpmpc/solve_mpc.py:43``
of the reference):   min 1/2 x'Px + q'x   s.t.   G x <= h.

These are the known-answer vectors that pin the solver half of the oracle and
of the CUDA kernels to results obtained by others -- the wheels the reference
calls (proxsuite, quadprog) are not installable here, but their documented
examples and the standard test-set optima are:

* ``quadprog_doc``  the example of ``solve.QP`` in the manual of the R package
  quadprog (B. Turlach, A. Weingessel: "quadprog: Functions to Solve Quadratic
  Programming Problems", CRAN, help page of solve.QP) -- the Goldfarb-Idnani
  code the ``quadprog`` backend of qpsolvers wraps.  Printed solution
  0.4761905 1.0476190 2.0952381, value -2.380952, Lagrangian 0 0.2380952
  2.0952381 (= 10/21, 22/21, 44/21; -50/21).
* ``qpsolvers_readme``  the example of the README of qpsolvers (S. Caron et
  al., "qpsolvers: Quadratic Programming Solvers in Python"), whose printed
  answer is [0.30769231, -0.69230769, 1.38461538] (= 4/13, -9/13, 18/13); its
  equality row is written as two inequalities.
* ``hs21``, ``hs35``, ``hs76``, ``hs118``, ``hs268``  W. Hock, K. Schittkowski,
  "Test Examples for Nonlinear Programming Codes", Lecture Notes in Economics
  and Mathematical Systems 187, Springer 1981, problems 21, 35, 76, 118, 268
  (268 from the 1987 sequel, K. Schittkowski, "More Test Examples ...", LNEMS
  282), with the solutions x* and values f* printed there; the same five are
  part of the convex QP test set of I. Maros and C. Meszaros, "A repository of
  convex quadratic programming problems", Optimization Methods and Software
  11 (1999), whose table lists the optimal values -9.99599999e+01,
  1.11111111e-01, -4.68181818e+00, 6.64820450e+02 and 5.73107049e-07 (the last
  is 0 up to the accuracy of the solver that produced the table).

Each entry: P, q, G, h, the published x*, the published objective value
``f`` (of 1/2 x'Px + q'x + ``const``) and ``x_tol``, the number of digits the
publication prints.
"""

import numpy as np


def _bounds(n, lo, hi):
    """Rows of G, h for lo <= x <= hi (None entries: no bound)."""
    G, h = [], []
    for i in range(n):
        if hi[i] is not None:
            r = np.zeros(n); r[i] = 1.0
            G.append(r); h.append(hi[i])
        if lo[i] is not None:
            r = np.zeros(n); r[i] = -1.0
            G.append(r); h.append(-lo[i])
    return G, h


def quadprog_doc():
    # min 1/2 x'Dx - d'x  s.t.  A'x >= b   (R: solve.QP(Dmat, dvec, Amat, bvec))
    A = np.array([[-4.0, -3.0, 0.0], [2.0, 1.0, 0.0], [0.0, -2.0, 1.0]])  # rows = constraints
    b = np.array([-8.0, 2.0, 0.0])
    return dict(name="quadprog_doc", P=np.eye(3), q=-np.array([0.0, 5.0, 0.0]), G=-A, h=-b, const=0.0,
                x=np.array([10.0, 22.0, 44.0]) / 21.0, f=-50.0 / 21.0, x_tol=5e-8,
                z=np.array([0.0, 5.0 / 21.0, 44.0 / 21.0]))


def qpsolvers_readme():
    M = np.array([[1.0, 2.0, 0.0], [-8.0, 3.0, 2.0], [0.0, 1.0, 1.0]])
    G = np.array([[1.0, 2.0, 1.0], [2.0, 0.0, 1.0], [-1.0, 2.0, -1.0]])
    h = np.array([3.0, 2.0, -2.0])
    A, b = np.array([[1.0, 1.0, 1.0]]), np.array([1.0])
    return dict(name="qpsolvers_readme", P=M.T @ M, q=np.array([3.0, 2.0, 3.0]) @ M,
                G=np.vstack([G, A, -A]), h=np.concatenate([h, b, -b]), const=0.0,
                x=np.array([4.0, -9.0, 18.0]) / 13.0, f=None, x_tol=5e-9, z=None)


def hs21():
    # min 0.01 x1^2 + x2^2 - 100;  10 x1 - x2 >= 10;  2 <= x1 <= 50;  -50 <= x2 <= 50
    G, h = _bounds(2, [2.0, -50.0], [50.0, 50.0])
    G.append(np.array([-10.0, 1.0])); h.append(-10.0)
    return dict(name="hs21", P=np.diag([0.02, 2.0]), q=np.zeros(2), G=np.array(G), h=np.array(h),
                const=-100.0, x=np.array([2.0, 0.0]), f=-99.96, x_tol=1e-9, z=None)


def hs35():
    # min 9 - 8x1 - 6x2 - 4x3 + 2x1^2 + 2x2^2 + x3^2 + 2x1x2 + 2x1x3;  x1 + x2 + 2x3 <= 3;  x >= 0
    G, h = _bounds(3, [0.0] * 3, [None] * 3)
    G.append(np.array([1.0, 1.0, 2.0])); h.append(3.0)
    P = np.array([[4.0, 2.0, 2.0], [2.0, 4.0, 0.0], [2.0, 0.0, 2.0]])
    return dict(name="hs35", P=P, q=np.array([-8.0, -6.0, -4.0]), G=np.array(G), h=np.array(h), const=9.0,
                x=np.array([4.0 / 3.0, 7.0 / 9.0, 4.0 / 9.0]), f=1.0 / 9.0, x_tol=1e-9, z=None)


def hs76():
    # min x1^2 + .5x2^2 + x3^2 + .5x4^2 - x1x3 + x3x4 - x1 - 3x2 + x3 - x4
    # x1 + 2x2 + x3 + x4 <= 5;  3x1 + x2 + 2x3 - x4 <= 4;  x2 + 4x3 >= 1.5;  x >= 0
    G, h = _bounds(4, [0.0] * 4, [None] * 4)
    G += [np.array([1.0, 2.0, 1.0, 1.0]), np.array([3.0, 1.0, 2.0, -1.0]), np.array([0.0, -1.0, -4.0, 0.0])]
    h += [5.0, 4.0, -1.5]
    P = np.array([[2.0, 0.0, -1.0, 0.0], [0.0, 1.0, 0.0, 0.0], [-1.0, 0.0, 2.0, 1.0], [0.0, 0.0, 1.0, 1.0]])
    return dict(name="hs76", P=P, q=np.array([-1.0, -3.0, 1.0, -1.0]), G=np.array(G), h=np.array(h), const=0.0,
                x=np.array([3.0, 23.0, 0.0, 6.0]) / 11.0, f=-103.0 / 22.0, x_tol=5e-8, z=None)


def hs118():
    n = 15
    lin = np.tile([2.3, 1.7, 2.2], 5)
    quad = np.tile([0.0001, 0.0001, 0.00015], 5)
    lo = [8.0, 43.0, 3.0] + [0.0] * 12
    hi = [21.0, 57.0, 16.0] + [90.0, 120.0, 60.0] * 4
    G, h = _bounds(n, lo, hi)
    width = [13.0, 14.0, 13.0]  # 0 <= x_{3j+i} - x_{3j+i-3} + 7 <= width_i
    for j in range(1, 5):
        for i in range(3):
            r = np.zeros(n); r[3 * j + i] = 1.0; r[3 * j + i - 3] = -1.0
            G.append(r.copy()); h.append(width[i] - 7.0)
            G.append(-r); h.append(7.0)
    for j, need in enumerate([60.0, 50.0, 70.0, 85.0, 100.0]):
        r = np.zeros(n); r[3 * j:3 * j + 3] = -1.0
        G.append(r); h.append(-need)
    x = np.array([8, 49, 3, 1, 56, 0, 1, 63, 6, 3, 70, 12, 5, 77, 18], dtype=float)
    return dict(name="hs118", P=np.diag(2.0 * quad), q=lin, G=np.array(G), h=np.array(h), const=0.0,
                x=x, f=664.82045, x_tol=1e-7, z=None)


def hs268():
    # min 14463 + x'Dx - 2 d'x subject to five linear inequalities; x* = (1, 2, -1, 3, -4), f* = 0
    D = np.array([[10197.0, -12454.0, -1013.0, 1948.0, 329.0], [-12454.0, 20909.0, -1733.0, -4914.0, -186.0],
                  [-1013.0, -1733.0, 1755.0, 1089.0, -174.0], [1948.0, -4914.0, 1089.0, 1515.0, -22.0],
                  [329.0, -186.0, -174.0, -22.0, 27.0]])
    d = np.array([-9170.0, 17099.0, -2271.0, -4336.0, -43.0])
    # a_i'x + b_i >= 0
    A = np.array([[-1.0, -1.0, -1.0, -1.0, -1.0], [10.0, 10.0, -3.0, 5.0, 4.0], [-8.0, 1.0, -2.0, -5.0, 3.0],
                  [8.0, -1.0, 2.0, 5.0, -3.0], [-4.0, -2.0, 3.0, -5.0, 1.0]])
    b = np.array([5.0, -20.0, 40.0, -11.0, 30.0])
    return dict(name="hs268", P=2.0 * D, q=-2.0 * d, G=-A, h=b, const=14463.0,
                x=np.array([1.0, 2.0, -1.0, 3.0, -4.0]), f=0.0, x_tol=1e-8, z=None)


ALL = (quadprog_doc, qpsolvers_readme, hs21, hs35, hs76, hs118, hs268)


def as_one_step_mpc(qp):
    """The QP as a one-step MPC problem (N = 1, nx = nu = n, nc = m), so that it can go through
    the condensing + solve path: with psi_1 = B, P = w_u I + w_t B'B and q = w_t B'(A x0 - goal)
    (``qpmpc/mpc_qp.py:99-105,139-149``); the k = 0 rows are G_0 = D_0, h_0 = e_0 (``:67-78``)
    because psi_0 = 0.  Choose w_u < lambda_min(P), B'B = P - w_u I, x0 = 0, goal = -B'^-1 q.
    """
    P, q, G, h = qp["P"], qp["q"], qp["G"], qp["h"]
    n, m = q.size, h.size
    w_u = 0.5 * float(np.linalg.eigvalsh(P)[0])
    B = np.linalg.cholesky(P - w_u * np.eye(n)).T  # B'B = P - w_u I
    goal = -np.linalg.solve(B.T, q)
    return dict(name=qp["name"], batch=1, N=1, nx=n, nu=n, nc=m, A=np.eye(n), B=B, C=None, D=G.copy(),
                e=h.copy(), x0=np.zeros((1, n)), goal=goal[None], targets=None, w_t=1.0, w_x=None,
                w_u=w_u, ltv=())
