"""ORACLE (test infrastructure, never the product path): float64 primal-dual
interior-point solver for

    minimise   1/2 x'Px + q'x
    subject to Gx + s = h,  s in K = R_+^l  x  (SOC_3)^nq,     Ax = b.

What it restates.  The reference solves its MPC problem with
``cp.Problem(...).solve(solver=self.solver)`` (reference
adacharge/adaptive_charging_optimization.py:315-318), default solver "ECOS"
(:37).  cvxpy and ECOS are third-party, unpinned (reference setup.py:24) and absent
from /root/reference and from this image, so the algorithm is restated from its
published description: ECOS (Domahidi, Chu, Boyd, ECC 2013) is a Mehrotra
predictor-corrector primal-dual interior-point method with Nesterov-Todd scaling
that factorises the sparse indefinite KKT system each iteration; the variant here
follows the infeasible-start cone-QP method of Vandenberghe, "The CVXOPT linear
and quadratic cone program solvers" (2010), which keeps a quadratic objective
instead of lifting it into a cone.  Infeasibility is decided by a phase-1 problem
(``phase1``) rather than by ECOS's self-dual embedding.

All second-order cones on this path have dimension 3 (a limit and the two
rectangular components of a phase current), so cone arithmetic is vectorised over
an (nq, 3) array.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as sla


class IPMResult:
    def __init__(self, status, x, s, z, y, iters, pcost, dcost, gap, pres, dres):
        self.status = status
        self.x, self.s, self.z, self.y = x, s, z, y
        self.iters = iters
        self.pcost, self.dcost, self.gap, self.pres, self.dres = pcost, dcost, gap, pres, dres


# ----------------------------------------------------------------- cone helpers
def _split(v, l):
    return v[:l], v[l:].reshape(-1, 3)


def _jdot(u, v):  # u'Jv for SOC blocks, J = diag(1,-1,-1)
    return u[:, 0] * v[:, 0] - u[:, 1] * v[:, 1] - u[:, 2] * v[:, 2]


def _max_step(v, dv, l):
    """Largest alpha in [0, inf) with v + alpha*dv in K (v strictly inside)."""
    vl, vq = _split(v, l)
    dl, dq = _split(dv, l)
    alpha = np.inf
    neg = dl < 0
    if neg.any():
        alpha = min(alpha, np.min(-vl[neg] / dl[neg]))
    if len(vq):
        # (v0+a d0)^2 - |v1+a d1|^2 >= 0  and v0 + a d0 >= 0
        a = _jdot(dq, dq)
        b = 2 * _jdot(vq, dq)
        c = _jdot(vq, vq)
        disc = b * b - 4 * a * c
        with np.errstate(divide="ignore", invalid="ignore"):
            sq = np.sqrt(np.maximum(disc, 0))
            r1 = (-b - sq) / (2 * a)
            r2 = (-b + sq) / (2 * a)
            lin = np.where(np.abs(b) > 0, -c / b, np.inf)
        cand = np.full(len(vq), np.inf)
        tiny = np.abs(a) < 1e-300
        for r in (r1, r2):
            ok = (~tiny) & (disc >= 0) & (r > 0)
            cand = np.where(ok, np.minimum(cand, r), cand)
        okl = tiny & (lin > 0)
        cand = np.where(okl, np.minimum(cand, lin), cand)
        neg0 = dq[:, 0] < 0
        cand = np.where(neg0, np.minimum(cand, -vq[:, 0] / np.where(neg0, dq[:, 0], -1.0)), cand)
        alpha = min(alpha, cand.min())
    return alpha


def _nt_scaling(s, z, l):
    """Nesterov-Todd scaling: returns (w_l, Wq (nq,3,3), lambda)."""
    sl, sq = _split(s, l)
    zl, zq = _split(z, l)
    wl = np.sqrt(sl / zl)
    lam_l = np.sqrt(sl * zl)
    nq = len(sq)
    if nq == 0:
        return wl, np.zeros((0, 3, 3)), lam_l
    sn = np.sqrt(_jdot(sq, sq))
    zn = np.sqrt(_jdot(zq, zq))
    sb = sq / sn[:, None]
    zb = zq / zn[:, None]
    gamma = np.sqrt((1 + np.einsum("ki,ki->k", sb, zb)) / 2)
    jz = zb * np.array([1.0, -1.0, -1.0])
    wb = (sb + jz) / (2 * gamma[:, None])
    v = wb.copy()
    v[:, 0] += 1
    v /= np.sqrt(2 * (wb[:, 0] + 1))[:, None]
    beta = np.sqrt(sn / zn)
    J = np.diag([1.0, -1.0, -1.0])
    W = beta[:, None, None] * (2 * np.einsum("ki,kj->kij", v, v) - J[None])
    lam_q = np.einsum("kij,kj->ki", W, zq)
    return wl, W, np.concatenate([lam_l, lam_q.ravel()])


def _jprod(u, v, l):
    ul, uq = _split(u, l)
    vl, vq = _split(v, l)
    out_l = ul * vl
    out_q = np.empty_like(uq)
    if len(uq):
        out_q[:, 0] = np.einsum("ki,ki->k", uq, vq)
        out_q[:, 1:] = uq[:, :1] * vq[:, 1:] + vq[:, :1] * uq[:, 1:]
    return np.concatenate([out_l, out_q.ravel()])


def _jdiv(lam, b, l):
    """Solve lam o x = b."""
    ll, lq = _split(lam, l)
    bl, bq = _split(b, l)
    out_l = bl / ll
    out_q = np.empty_like(bq)
    if len(bq):
        det = _jdot(lq, lq)
        x0 = (lq[:, 0] * bq[:, 0] - np.einsum("ki,ki->k", lq[:, 1:], bq[:, 1:])) / det
        out_q[:, 0] = x0
        out_q[:, 1:] = (bq[:, 1:] - x0[:, None] * lq[:, 1:]) / lq[:, :1]
    return np.concatenate([out_l, out_q.ravel()])


def _apply_W(wl, W, v, l, inverse=False):
    vl, vq = _split(v, l)
    out_l = vl / wl if inverse else vl * wl
    if len(vq):
        M = np.linalg.inv(W) if inverse else W
        out_q = np.einsum("kij,kj->ki", M, vq).ravel()
    else:
        out_q = np.zeros(0)
    return np.concatenate([out_l, out_q])


def _identity(l, nq):
    e = np.zeros(l + 3 * nq)
    e[:l] = 1
    e[l::3] = 1
    return e


# ------------------------------------------------------------------------ solver
def solve(P, q, G, h, l, nq, A=None, b=None, max_iter=100, feastol=1e-9, abstol=1e-9,
          reltol=1e-9, verbose=False):
    """Returns IPMResult with status 'optimal' or 'unknown'."""
    n = len(q)
    m = l + 3 * nq
    P = sp.csc_matrix(P) if P is not None else sp.csc_matrix((n, n))
    G = sp.csc_matrix(G)
    if A is None:
        A = sp.csc_matrix((0, n))
        b = np.zeros(0)
    A = sp.csc_matrix(A)
    p = A.shape[0]
    e = _identity(l, nq)
    Gt, At = G.T.tocsc(), A.T.tocsc()

    cone_r = (l + 3 * np.repeat(np.arange(nq), 9) + np.tile(np.repeat(np.arange(3), 3), nq)) if nq else np.zeros(0, int)
    cone_c = (l + 3 * np.repeat(np.arange(nq), 9) + np.tile(np.tile(np.arange(3), 3), nq)) if nq else np.zeros(0, int)
    reg_x = 1e-11 * sp.identity(n, format="csc")
    reg_y = -1e-11 * sp.identity(p, format="csc")

    def kkt_factor(wl, W):
        # [[P, A', G'], [A, -eps, 0], [G, 0, -W'W]]; symmetric quasi-definite, so a
        # fill-reducing ordering on the pattern of K + K' is the right one.
        rows = np.concatenate([np.arange(l), cone_r])
        cols = np.concatenate([np.arange(l), cone_c])
        vals = np.concatenate([wl * wl, np.einsum("kij,kjl->kil", W, W).ravel()])
        blk = sp.csc_matrix((vals, (rows, cols)), shape=(m, m))
        K = sp.bmat([[P + reg_x, At, Gt], [A, reg_y, None], [G, None, -blk]], format="csc")
        return sla.splu(K, permc_spec="MMD_AT_PLUS_A")

    def kkt_solve(lu, wl, W, lam, bx, by, bz, bs):
        # bs is the rhs of the complementarity equation lam o (W dz + W^-T ds) = bs
        t = _jdiv(lam, bs, l)
        Wt = _apply_W(wl, W, t, l)  # W' = W (symmetric)
        rhs = np.concatenate([bx, by, bz - Wt])
        sol = lu.solve(rhs)
        dx, dy, dz = sol[:n], sol[n : n + p], sol[n + p :]
        ds = Wt - _apply_W(wl, W, _apply_W(wl, W, dz, l), l)
        return dx, dy, dz, ds

    # ---- initial point (CVXOPT coneqp): W = I
    one_l = np.ones(l)
    I3 = np.tile(np.eye(3), (nq, 1, 1))
    lu = kkt_factor(one_l, I3)
    sol = lu.solve(np.concatenate([-q, b, h]))
    x, y, z = sol[:n], sol[n : n + p], sol[n + p :]
    s = -z.copy()

    def shift(v):
        vl, vq = _split(v, l)
        t = -np.inf
        if l:
            t = max(t, (-vl).max())
        if nq:
            t = max(t, (np.hypot(vq[:, 1], vq[:, 2]) - vq[:, 0]).max())
        if t >= -1e-8 * max(1.0, np.linalg.norm(v)):
            return v + (1 + t) * e
        return v

    s, z = shift(s), shift(z)

    resx0 = max(1.0, np.linalg.norm(q))
    resy0 = max(1.0, np.linalg.norm(b))
    resz0 = max(1.0, np.linalg.norm(h))
    status = "unknown"
    it = 0
    pcost = dcost = gap = pres = dres = np.nan
    for it in range(max_iter + 1):
        Px = P @ x
        rx = Px + q + At @ y + Gt @ z
        ry = A @ x - b
        rz = G @ x + s - h
        gap = float(s @ z)
        f0 = 0.5 * x @ Px + q @ x
        pcost = f0
        dcost = f0 + y @ ry + z @ rz - gap
        pres = max(np.linalg.norm(ry) / resy0, np.linalg.norm(rz) / resz0)
        dres = np.linalg.norm(rx) / resx0
        if pcost < 0:
            relgap = gap / -pcost
        elif dcost > 0:
            relgap = gap / dcost
        else:
            relgap = np.inf
        if verbose:
            print(f"{it:3d} pcost {pcost: .8e} dcost {dcost: .8e} gap {gap:.2e} pres {pres:.2e} dres {dres:.2e}")
        if pres <= feastol and dres <= feastol and (gap <= abstol or relgap <= reltol):
            status = "optimal"
            break
        if it == max_iter:
            break
        wl, W, lam = _nt_scaling(s, z, l)
        try:
            lu = kkt_factor(wl, W)
        except RuntimeError:
            break
        mu = gap / (l + nq)
        lam2 = _jprod(lam, lam, l)
        # affine direction
        dxa, dya, dza, dsa = kkt_solve(lu, wl, W, lam, -rx, -ry, -rz, -lam2)
        a_aff = min(1.0, _max_step(s, dsa, l), _max_step(z, dza, l))
        sigma = (1 - a_aff) ** 3
        # combined direction
        ws = _apply_W(wl, W, dsa, l, inverse=True)  # W^-T ds
        wz = _apply_W(wl, W, dza, l)
        bs = -lam2 - _jprod(ws, wz, l) + sigma * mu * e
        dx, dy, dz, ds = kkt_solve(lu, wl, W, lam, -(1 - sigma) * rx, -(1 - sigma) * ry, -(1 - sigma) * rz, bs)
        alpha = min(1.0, 0.99 * min(_max_step(s, ds, l), _max_step(z, dz, l)))
        if not np.isfinite(alpha) or alpha < 1e-14:
            break
        x = x + alpha * dx
        y = y + alpha * dy
        z = z + alpha * dz
        s = s + alpha * ds
    return IPMResult(status, x, s, z, y, it, pcost, dcost, gap, pres, dres)


def phase1(G, h, l, nq, A=None, b=None, groups=None, n_groups=1, **kw):
    """min sum(tau)  s.t.  Gx + s = h + E tau,  tau >= 0,  Ax = b, where column g of E
    is the cone identity restricted to the rows of relaxation group g (``groups[r]`` for
    orthant row r < l and for cone r - l >= 0).  One group = the classic single-tau
    phase 1; per-period / per-session groups keep the KKT system sparse.
    Returns (sum tau*, IPMResult)."""
    n = G.shape[1]
    G = sp.csc_matrix(G)
    if groups is None:
        groups = np.zeros(l + nq, dtype=int)
    gl, gq = groups[:l], groups[l:]
    El = sp.csc_matrix((np.ones(l), (np.arange(l), gl)), shape=(l, n_groups))
    Eq = sp.csc_matrix((np.ones(nq), (3 * np.arange(nq), gq)), shape=(3 * nq, n_groups))
    top = sp.hstack([G[:l], -El])
    tau_rows = sp.hstack([sp.csc_matrix((n_groups, n)), -sp.identity(n_groups)])
    bot = sp.hstack([G[l:], -Eq])
    G1 = sp.vstack([top, tau_rows, bot]).tocsc()
    h1 = np.concatenate([h[:l], np.zeros(n_groups), h[l:]])
    q1 = np.concatenate([np.zeros(n), np.ones(n_groups)])
    A1 = None
    if A is not None and A.shape[0]:
        A1 = sp.hstack([sp.csc_matrix(A), sp.csc_matrix((A.shape[0], n_groups))]).tocsc()
    P1 = sp.diags(np.concatenate([np.full(n, 1e-9), np.zeros(n_groups)]))
    res = solve(P1, q1, G1, h1, l + n_groups, nq, A1, b, **kw)
    return float(res.x[n:].sum()), res
