"""NumPy model of the iteration the CUDA kernel runs.  TEST INFRASTRUCTURE ONLY.

Batched (vectorised over instances) Mehrotra predictor-corrector interior
point for   min 1/2 u'Pu + q'u  s.t.  G u <= h   followed by an active-set
polish, written to mirror ``qpmpc_b200/csrc/mpc_kernels.cuh`` step for step so
that an algorithmic disagreement between the kernel and the exact active-set
oracle (``mpc_oracle.c``) can be told apart from an indexing bug.  It replaces
the third-party solve at ``qpmpc/solve_mpc.py:43`` (reference tree) in tests
that cross-check the Goldfarb-Idnani oracle with an independent method.

Newton step (slack form  G u + s = h, s, z >= 0,  W = z / s):
    (P + G'WG) du = -r_d - G'(W r_p - r_c / s)
    ds = -r_p - G du ;  dz = -(r_c + z ds) / s
with r_d = P u + q + G'z, r_p = G u + s - h, r_c = s z (predictor) or
s z + ds_a dz_a - sigma mu (corrector).  One step length for primal and dual,
so r_d and r_p shrink by exactly (1 - alpha).

Polish (a primal-dual active-set finish): rows with z > s are taken as the
active set A; the equality QP on A is solved by a few proximal multiplier steps
in residual form, which reuse the same factorisation code,
    r1 = P u + q + G_A' lam,  r2 = G_A u - h_A,
    (P + G_A' G_A / delta) du = -(r1 + G_A' r2 / delta),  dlam = (r2 + G_A du) / delta
(the residuals are evaluated exactly, so the ill-conditioning of the proximal
matrix only slows convergence, it does not limit the accuracy).  The polished
point is accepted if it is primal feasible, its multipliers are non-negative
and its stationarity residual is small; otherwise A is updated -- rows with a
negative multiplier leave, violated rows enter -- and the polish is repeated,
up to ``polish_rounds`` times.  An accepted point is the exact solution on its
active set, i.e. what the active-set oracle returns.
"""

import numpy as np


def _solve(L, b):
    y = np.linalg.solve(L, b[..., None])
    return np.linalg.solve(np.swapaxes(L, -1, -2), y)[..., 0]


def pdip_batch(P, q, G, h, max_iter=40, tol=1e-9, polish=True, polish_steps=8,
               delta=1e-7, polish_rounds=8):
    """Solve a batch of QPs.  P [B,n,n], q [B,n], G [B,m,n], h [B,m].

    Returns dict(U, z, status, iters): status 0 solved, 1 max_iter, 2 numerical.
    """
    P, q, G, h = (np.asarray(a, dtype=np.float64) for a in (P, q, G, h))
    B, m, n = G.shape
    u = np.zeros((B, n))
    # start: u = 0, s = max(h, 1), z = 1 / s (every complementarity product starts at 1: a row
    # with a huge bound -- "no bound" constants, padding -- does not inflate mu)
    s = np.maximum(h - np.einsum("bmn,bn->bm", G, u), 1.0)
    z = 1.0 / s
    status = np.full(B, 1, dtype=np.int32)
    iters = np.zeros(B, dtype=np.int32)
    active = np.ones(B, dtype=bool)
    hrow = np.maximum(1.0, np.abs(h))  # per-row scale of the primal tests
    qscale = np.maximum(1.0, np.abs(q).max(axis=1))
    for it in range(max_iter):
        Pu = np.einsum("bij,bj->bi", P, u)
        Gtz = np.einsum("bmn,bm->bn", G, z)
        Gu = np.einsum("bmn,bn->bm", G, u)
        r_d = Pu + q + Gtz
        r_p = Gu + s - h
        mu = (s * z).sum(axis=1) / max(m, 1)
        # relative criteria: each residual against the size of the terms it is the sum of
        # (their rounding noise is the floor it can reach), the gap against the objective
        dscale = np.maximum(qscale, np.maximum(np.abs(Pu), np.abs(Gtz)).max(axis=1))
        # (the primal residual row by row: one huge bound must not relax the test of the others)
        prow = np.maximum(hrow, np.maximum(np.abs(Gu), s))
        obj = np.abs(np.einsum("bi,bi->b", u, 0.5 * Pu + q))
        done = (
            (np.abs(r_d).max(axis=1) <= tol * dscale)
            & (np.abs(r_p) <= tol * prow).all(axis=1)
            & (mu <= tol * (1.0 + obj))
        )
        newly = active & done
        status[newly] = 0
        active &= ~done
        if not active.any():
            break
        iters[active] += 1
        W = z / s
        H = P + np.einsum("bmi,bm,bmj->bij", G, W, G)
        L, spd = _cholesky_each(H)
        lost = active & ~spd
        status[lost] = 3  # H lost positive definiteness (numerically): the polish may still rescue it
        active &= spd
        if not active.any():
            break
        # predictor
        rhs = -r_d - np.einsum("bmn,bm->bn", G, W * r_p - z)
        du = _solve(L, rhs)
        ds = -r_p - np.einsum("bmn,bn->bm", G, du)
        dz = -(s * z + z * ds) / s
        a_aff = _step(s, z, ds, dz, 1.0)
        mu_aff = ((s + a_aff[:, None] * ds) * (z + a_aff[:, None] * dz)).sum(axis=1) / m
        sigma = (mu_aff / mu) ** 3
        # corrector
        r_c = s * z + ds * dz - (sigma * mu)[:, None]
        rhs = -r_d - np.einsum("bmn,bm->bn", G, W * r_p - r_c / s)
        du = _solve(L, rhs)
        ds = -r_p - np.einsum("bmn,bn->bm", G, du)
        dz = -(r_c + z * ds) / s
        alpha = _step(s, z, ds, dz, 0.99)
        alpha = np.where(active, alpha, 0.0)[:, None]
        u = u + alpha * du
        s = s + alpha * ds
        z = z + alpha * dz
    if polish:
        # instances the iteration gave up on (cap, lost definiteness) are polished too: an accepted
        # point is a KKT-certified solution wherever the iterate came from; they are held to the
        # absolute form of the acceptance test (their iterate may be huge)
        up, zp, ok = _polish(P, q, G, h, u, s, z, polish_steps, delta, polish_rounds, strict=status != 0)
        status[ok] = 0
        u = np.where(ok[:, None], up, u)
        z = np.where(ok[:, None], zp, z)
    else:
        ok = np.zeros(B, dtype=bool)
    return dict(U=u, z=z, status=status, iters=iters, polished=ok)


def _step(s, z, ds, dz, frac):
    """Largest alpha in (0, 1] keeping s + alpha ds, z + alpha dz positive."""
    with np.errstate(divide="ignore", invalid="ignore"):
        a_s = np.where(ds < 0, -s / ds, np.inf).min(axis=1)
        a_z = np.where(dz < 0, -z / dz, np.inf).min(axis=1)
    return np.minimum(1.0, frac * np.minimum(a_s, a_z))


def _cholesky_each(H):
    """Batched Cholesky that survives a failing member: (L, ok [B])."""
    try:
        return np.linalg.cholesky(H), np.ones(H.shape[0], dtype=bool)
    except np.linalg.LinAlgError:
        L, ok = np.zeros_like(H), np.zeros(H.shape[0], dtype=bool)
        for b in range(H.shape[0]):
            try:
                L[b], ok[b] = np.linalg.cholesky(H[b]), True
            except np.linalg.LinAlgError:
                L[b] = np.eye(H.shape[1])
        return L, ok


def _polish(P, q, G, h, u, s, z, steps, delta, rounds, eps=1e-9, strict=None):
    B = q.shape[0]
    strict = np.zeros(B, dtype=bool) if strict is None else strict
    act = z > s
    lam = np.where(act, z, 0.0)
    up = u.copy()
    u_out, z_out = u.copy(), z.copy()
    accepted = np.zeros(B, dtype=bool)
    hrow = np.maximum(1.0, np.abs(h))
    qscale = np.maximum(1.0, np.abs(q).max(axis=1))
    pmin = np.linalg.eigvalsh(P)[:, 0]  # the kernel uses w_u <= lambda_min(P)
    for _ in range(rounds):
        if accepted.all():
            break
        H = P + np.einsum("bmi,bm,bmj->bij", G, act / delta, G)
        L, spd = _cholesky_each(H)
        lam = np.where(act, lam, 0.0)
        frozen = accepted.copy()
        for step in range(steps + 1):
            # residuals first; steps move (up, lam) until the stationarity residual is as small as
            # the parity bar needs (|r1| <= 1e-9 lambda_min(P) |u|) or as rounding allows
            Pu = np.einsum("bij,bj->bi", P, up)
            Gtl = np.einsum("bmn,bm->bn", G, lam)
            Gu = np.einsum("bmn,bn->bm", G, up)
            r1 = Pu + q + Gtl
            viol = Gu - h
            r2 = np.where(act, viol, 0.0)
            pscale = np.where(strict[:, None], hrow, np.maximum(hrow, np.abs(Gu)))  # row by row
            dscale = np.where(strict, qscale, np.maximum(qscale, np.maximum(np.abs(Pu), np.abs(Gtl)).max(axis=1)))
            uscale = np.maximum(1.0, np.abs(up).max(axis=1))
            rd_tol = np.maximum(1e-9 * pmin * uscale, 64 * 2.3e-16 * dscale)
            with np.errstate(invalid="ignore"):
                tight = (np.abs(r1).max(axis=1) <= rd_tol) & (np.abs(r2) <= eps * pscale).all(axis=1)
            frozen |= tight
            if step == steps or frozen.all():
                break
            du = _solve(L, -(r1 + np.einsum("bmn,bm->bn", G, r2 / delta)))
            du = np.where(frozen[:, None], 0.0, du)
            lam = lam + np.where(act & ~frozen[:, None], (r2 + np.einsum("bmn,bn->bm", G, du)) / delta, 0.0)
            up = up + du
        zscale = np.maximum(1.0, np.abs(lam).max(axis=1, initial=0.0))[:, None]
        with np.errstate(invalid="ignore"):
            finite = np.isfinite(up).all(axis=1) & np.isfinite(lam).all(axis=1)
            ok = (
                spd & finite
                & (viol <= eps * pscale).all(axis=1)
                & (np.abs(np.where(act, viol, 0.0)) <= eps * pscale).all(axis=1)
                # (a multiplier of -e moves u by up to e |g| / lambda_min(P): held to the same bar)
                & (lam >= -np.maximum(1e-9 * pmin * uscale, 64 * 2.3e-16 * zscale[:, 0])[:, None]).all(axis=1)
                & (np.abs(r1).max(axis=1) <= rd_tol)
            )
        new = ok & ~accepted
        u_out[new], z_out[new] = up[new], lam[new]
        accepted |= ok
        # correct the guess by ONE row: the most violated row enters; if nothing is violated the
        # most negative multiplier leaves (changing many rows at once can cycle)
        vrel = np.where(act, -np.inf, viol / pscale)
        worst = vrel.argmax(axis=1)
        has_viol = vrel.max(axis=1) > eps
        lneg = np.where(act, lam, np.inf)
        drop = lneg.argmin(axis=1)
        rows = np.arange(B)
        act = act.copy()
        act[rows[has_viol], worst[has_viol]] = True
        sel = ~has_viol & (lneg.min(axis=1) < 0.0)
        act[rows[sel], drop[sel]] = False
    return u_out, z_out, accepted
