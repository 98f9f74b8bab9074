"""ORACLE (test infrastructure, never the product path): CPU float64 restatement of
the reference's per-step MPC problem, reference
adacharge/adaptive_charging_optimization.py ("aco.py"):

* ``bounds``            <- charging_rate_bounds          aco.py:61-79
* energy rows           <- energy_constraints            aco.py:104-124
* SOC / LINEAR rows     <- infrastructure_constraints    aco.py:145-179
* peak rows             <- peak_constraint               aco.py:196-198
* ``objective_terms``   <- build_objective + objective functions  aco.py:200-218, 336-408
* ``solve_mpc``         <- build_problem + solve         aco.py:243-247, 310-321

The convex program is handed to ``oracle.conic_ipm`` (the restated ECOS-class
interior-point method) and, where it is an LP, cross-checked against HiGHS
(``solve_lp_highs``).  Parity status: the reference's own tests pin no rate matrix
or objective value at this boundary, and cvxpy/ECOS cannot run in
this image, so this oracle is pinned by (i) every property the reference's solver
tests assert (tests/test_oracle_mpc.py restates all scenarios of
tests/test_adaptive_charging_optimization.py), (ii) analytic unique optima of
those scenarios, (iii) HiGHS agreement on LP-representable cases, (iv) agreement with an
independent SLSQP solve on mixed-phase second-order-cone cases.  No output of the
reference solver itself is available: *solver parity is pinned to the reference's
tests and formulation, not to reference-produced vectors*.

Objective components are identified by the reference function names; the
coefficient/kwargs semantics are those of ObjectiveComponent (aco.py:12-15) with
component kwargs overriding caller kwargs (aco.py:203-217).
``non_completion_penalty`` does not exist in the reference;
it is defined by this project as  -sum_s |remaining_demand_s - E_s(R)|  (kWh)
(default, ``norm=1``; linear under the energy rows) or the sum of squares with ``norm=2``.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np
import scipy.sparse as sp

from . import conic_ipm


class OracleInfeasible(Exception):
    pass


def _name(f):
    return f if isinstance(f, str) else f.__name__


def horizon(sessions) -> int:
    return max(s.arrival_offset + s.remaining_time for s in sessions)  # aco.py:243-245


def bounds(sessions, station_ids: Sequence[str], T: int) -> Tuple[np.ndarray, np.ndarray]:
    N = len(station_ids)
    lb, ub = np.zeros((N, T)), np.zeros((N, T))
    ids = list(station_ids)
    for s in sessions:
        i = ids.index(s.station_id)
        a, e = s.arrival_offset, s.arrival_offset + s.remaining_time
        lb[i, a:e] = s.min_rates
        ub[i, a:e] = s.max_rates
    m = ub < lb
    ub[m] = lb[m]  # aco.py:75
    return lb, ub


def soc_rows(infra):
    """(M, 2, N): a_j = [v cos(phi); v sin(phi)]  (aco.py:156-158, utils.py:6-8)."""
    ph = np.deg2rad(infra.phases)
    A = np.asarray(infra.constraint_matrix, dtype=float)
    return np.stack([A * np.cos(ph), A * np.sin(ph)], axis=1)


def has_infrastructure(infra) -> bool:
    cm = infra.constraint_matrix
    return not (cm is None or np.asarray(cm).shape == (0, 0))  # aco.py:146-150


def objective_terms(objective, infra, interface, T, sessions=None, prev_peak=0):
    """Minimisation-form pieces of  -sum_c coef_c f_c(R):

    lin (N,T), diag_q scalar (adds diag_q*sum R^2), agg list of (gamma, ext[T])
    (adds gamma*sum_t (u_t+ext_t)^2, u = k'R), peaks list of (weight, p0)
    (adds weight*max(max_t u_t, p0)), ncp list of (weight, norm), const.
    """
    N = len(infra.station_ids)
    k = np.asarray(infra.voltages, dtype=float) / 1e3  # kW per A (aco.py:338-339)
    hrs = interface.period / 60.0  # aco.py:349
    lin = np.zeros((N, T))
    out = {"lin": lin, "diag_q": 0.0, "agg": [], "peaks": [], "ncp": [], "const": 0.0}
    for comp in objective:
        fn, coef = _name(comp[0]), comp[1]
        kw = dict(prev_peak=prev_peak)
        kw.update(comp[2] if len(comp) > 2 and comp[2] else {})
        if fn == "quick_charge":  # aco.py:363-371
            c = np.array([(T - t) / T for t in range(T)])
            lin -= coef * c[None, :]
        elif fn == "equal_share":  # aco.py:374-375
            out["diag_q"] += coef
        elif fn == "tou_energy_cost":  # aco.py:378-380
            prices = np.asarray(interface.get_prices(T), dtype=float)
            lin += coef * (k * hrs)[:, None] * prices[None, :]
        elif fn == "total_energy":  # aco.py:383-384
            lin -= coef * (k * hrs)[:, None]
        elif fn in ("peak", "demand_charge"):  # aco.py:387-400
            prev = interface.get_prev_peak() * infra.voltages[0] / 1000  # aco.py:390
            base = kw.get("baseline_peak", 0)
            p0 = max(prev, base) if base > 0 else prev
            w = -coef if fn == "peak" else coef * interface.get_demand_charge()
            if w < 0:
                raise ValueError(f"{fn} with this sign is not concave (cvxpy would raise a DCP error)")
            out["peaks"].append((w, p0))
        elif fn == "load_flattening":  # aco.py:403-408
            ext = kw.get("external_signal")
            ext = np.zeros(T) if ext is None else np.asarray(ext, dtype=float)[:T]
            if coef < 0:
                raise ValueError("load_flattening with negative coefficient is not concave")
            out["agg"].append((coef, ext))
        elif fn == "non_completion_penalty":  # project-defined, see module docstring
            if coef < 0:
                raise ValueError("non_completion_penalty with negative coefficient is not concave")
            out["ncp"].append((coef, int(kw.get("norm", 1))))
        elif fn == "linear":  # test hook, not a reference component: f = sum(weights * R), weights (N, >=T); the form a
            # user-written linear ObjectiveComponent.function (aco.py:200-218) canonicalises to
            lin -= coef * np.asarray(kw["weights"], dtype=float)[:, :T]
        else:
            raise ValueError(f"unknown objective component {fn}")
    if out["diag_q"] < 0:
        raise ValueError("equal_share with negative total coefficient is not concave")
    return out


def session_rows(sessions, infra, period):
    """[(i, start, stop, w_i, remaining_demand)] per session (aco.py:105-123)."""
    rows = []
    for s in sessions:
        i = infra.get_station_index(s.station_id)
        w = infra.voltages[i] * period / 1e3 / 60
        rows.append((i, s.arrival_offset, s.arrival_offset + s.remaining_time, float(w), float(s.remaining_demand)))
    return rows


def evaluate_objective(rates, objective, infra, interface, sessions=None, prev_peak=0) -> float:
    """Value of the reference's *maximised* objective sum_c coef_c f_c(rates)."""
    R = np.asarray(rates, dtype=float)
    T = R.shape[1]
    t = objective_terms(objective, infra, interface, T, sessions, prev_peak)
    k = np.asarray(infra.voltages, dtype=float) / 1e3
    u = k @ R
    val = (t["lin"] * R).sum() + t["diag_q"] * (R * R).sum()
    for g, ext in t["agg"]:
        val += g * ((u + ext) ** 2).sum()
    for w, p0 in t["peaks"]:
        val += w * max(u.max(), p0)
    if t["ncp"]:
        rows = session_rows(sessions if sessions is not None else interface.active_sessions(), infra, interface.period)
        unmet = np.array([e - w * R[i, a:b].sum() for (i, a, b, w, e) in rows])
        for wt, nrm in t["ncp"]:
            val += wt * (np.abs(unmet).sum() if nrm == 1 else (unmet**2).sum())
    return -float(val)


def violations(rates, sessions, infra, interface, constraint_type="SOC", peak_limit=None,
               enforce_energy_equality=False):
    """Max violation of each constraint family; infrastructure and peak relative to
    the limit, bounds in A, energy in kWh."""
    R = np.asarray(rates, dtype=float)
    T = R.shape[1]
    lb, ub = bounds(sessions, infra.station_ids, T)
    out = {"lb": float(np.max(lb - R)), "ub": float(np.max(R - ub))}
    en = 0.0
    for (i, a, b, w, e) in session_rows(sessions, infra, interface.period):
        d = w * R[i, a:b].sum() - e
        en = max(en, abs(d) if enforce_energy_equality else d)
    out["energy"] = float(en)
    inf = 0.0
    if has_infrastructure(infra):
        lim = np.asarray(infra.constraint_limits, dtype=float)
        if constraint_type == "SOC":
            a = soc_rows(infra)
            cur = np.sqrt((a[:, 0] @ R) ** 2 + (a[:, 1] @ R) ** 2)
        else:
            cur = np.abs(np.asarray(infra.constraint_matrix, dtype=float)) @ R
        inf = float(np.max((cur - lim[:, None]) / lim[:, None]))
    out["infrastructure_rel"] = inf
    if peak_limit is not None:
        pl = np.broadcast_to(np.asarray(peak_limit, dtype=float), (T,))
        out["peak_rel"] = float(np.max((R.sum(axis=0) - pl) / np.maximum(pl, 1e-12)))
    return out


# ---------------------------------------------------------------------- canonical form
def canonicalize(objective, sessions, infra, interface, constraint_type="SOC",
                 enforce_energy_equality=False, peak_limit=None, prev_peak=0):
    """Build (P, q, G, h, l, nq, A, b, meta) over the free variables.

    Variables fixed by lb == ub are eliminated (the interior-point method needs a
    strict interior); auxiliary variables: one epigraph variable per peak term, one
    unmet-energy variable per session for each non_completion_penalty term.
    """
    T = horizon(sessions)
    N = len(infra.station_ids)
    lb, ub = bounds(sessions, infra.station_ids, T)
    terms = objective_terms(objective, infra, interface, T, sessions, prev_peak)
    k = np.asarray(infra.voltages, dtype=float) / 1e3
    free = (ub - lb) > 0
    nfree = int(free.sum())
    col = -np.ones((N, T), dtype=int)
    col[free] = np.arange(nfree)
    x0 = np.where(free, 0.0, lb)  # fixed part
    rows_s = session_rows(sessions, infra, interface.period)
    n_pk = len(terms["peaks"])
    n_ncp = len(terms["ncp"]) * len(rows_s)
    n = nfree + n_pk + n_ncp
    fi, ft = np.nonzero(free)

    # objective
    q = np.zeros(n)
    q[:nfree] = terms["lin"][free]
    Pd = np.zeros(n)
    Pd[:nfree] = 2 * terms["diag_q"]
    const = (terms["lin"] * x0).sum() + terms["diag_q"] * (x0**2).sum()
    P = sp.diags(Pd).tocsc()
    u0 = k @ x0
    if terms["agg"]:
        # gamma * sum_t (k'x_t + u0_t + ext_t)^2 : per period rank-1
        U = sp.csr_matrix((k[fi], (ft, np.arange(nfree))), shape=(T, n))
        for g, ext in terms["agg"]:
            P = P + 2 * g * (U.T @ U)
            q += 2 * g * (U.T @ (u0 + ext))
            const += g * ((u0 + ext) ** 2).sum()
    Gl, hl, gl = [], [], []  # orthant rows, rhs, phase-1 relaxation group of each row
    # groups: 0..T-1 = period t, T.. = session index (rows owned by one session)
    owner = -np.ones((N, T), dtype=int)
    for r, (i, a, b_, w, e) in enumerate(rows_s):
        owner[i, a:b_] = r
    # bounds on free vars
    I = sp.identity(nfree, format="csr")
    pad = sp.csr_matrix((nfree, n - nfree))
    Gl.append(sp.hstack([-I, pad])); hl.append(-lb[free]); gl.append(T + owner[free])
    Gl.append(sp.hstack([I, pad])); hl.append(ub[free]); gl.append(T + owner[free])
    # energy
    Ae, be = [], []
    for (i, a, b_, w, e) in rows_s:
        cols = col[i, a:b_]
        cols = cols[cols >= 0]
        row = sp.csr_matrix((np.full(len(cols), w), (np.zeros(len(cols), dtype=int), cols)), shape=(1, n))
        rhs = e - w * x0[i, a:b_].sum()
        if enforce_energy_equality:
            Ae.append(row); be.append(rhs)
        else:
            Gl.append(row); hl.append([rhs]); gl.append([T + len(gl) - 2])
    # peak limit
    if peak_limit is not None:
        pl = np.broadcast_to(np.asarray(peak_limit, dtype=float), (T,))
        Srow = sp.csr_matrix((np.ones(nfree), (ft, np.arange(nfree))), shape=(T, n))
        Gl.append(Srow); hl.append(pl - x0.sum(axis=0)); gl.append(np.arange(T))
    # peak epigraphs: u_t - p <= 0 ; -p <= -p0
    for j, (w, p0) in enumerate(terms["peaks"]):
        pc = nfree + j
        q[pc] += w
        U = sp.csr_matrix((k[fi], (ft, np.arange(nfree))), shape=(T, n)).tolil()
        U[:, pc] = -1.0
        Gl.append(U.tocsr()); hl.append(-u0); gl.append(np.arange(T))
        Gl.append(sp.csr_matrix(([-1.0], ([0], [pc])), shape=(1, n))); hl.append([-p0]); gl.append([0])
    # non-completion: m_s = e_s - w sum R  (equality), penalty on m_s
    for j, (wt, nrm) in enumerate(terms["ncp"]):
        for r, (i, a, b_, w, e) in enumerate(rows_s):
            mc = nfree + n_pk + j * len(rows_s) + r
            cols = col[i, a:b_]
            cols = cols[cols >= 0]
            row = sp.csr_matrix(
                (np.concatenate([np.full(len(cols), w), [1.0]]),
                 (np.zeros(len(cols) + 1, dtype=int), np.concatenate([cols, [mc]]))), shape=(1, n))
            Ae.append(row); be.append(e - w * x0[i, a:b_].sum())
            if nrm == 1:
                # |m| with m >= 0 guaranteed only under the inequality energy rows; use epigraph m <= |m| via two rows
                # minimise wt*a, a >= m, a >= -m  -> reuse m directly when inequality rows keep m >= 0
                q[mc] += wt if not enforce_energy_equality else 0.0
            else:
                P = P + sp.csc_matrix(([2 * wt], ([mc], [mc])), shape=(n, n))
    # infrastructure
    Gq, hq = [], []
    if has_infrastructure(infra):
        lim = np.asarray(infra.constraint_limits, dtype=float)
        M = len(lim)
        if constraint_type == "SOC":
            if infra.phases is None:
                raise ValueError("phases is required when using SOC infrastructure constraints.")
            a = soc_rows(infra)
            for j in range(M):
                # cone (lim_j ; a_j0 . x_t ; a_j1 . x_t), one per t
                c0, c1 = a[j, 0], a[j, 1]
                base0, base1 = c0 @ x0, c1 @ x0
                nzmask = (np.abs(c0[fi]) + np.abs(c1[fi])) > 0
                idx = np.nonzero(nzmask)[0]
                r0 = sp.csr_matrix((-c0[fi[idx]], (3 * ft[idx] + 1, idx)), shape=(3 * T, n))
                r1 = sp.csr_matrix((-c1[fi[idx]], (3 * ft[idx] + 2, idx)), shape=(3 * T, n))
                Gq.append(r0 + r1)
                hh = np.zeros(3 * T)
                hh[0::3] = lim[j]
                hh[1::3] = base0
                hh[2::3] = base1
                hq.append(hh); gl.append(np.arange(T))
        elif constraint_type == "LINEAR":
            Aabs = np.abs(np.asarray(infra.constraint_matrix, dtype=float))
            for j in range(M):
                idx = np.nonzero(Aabs[j, fi] > 0)[0]
                Gl.append(sp.csr_matrix((Aabs[j, fi[idx]], (ft[idx], idx)), shape=(T, n)))
                hl.append(lim[j] - Aabs[j] @ x0); gl.append(np.arange(T))
        else:
            raise ValueError(
                "Invalid infrastructure constraint type: {0}. Valid options are SOC or AFFINE.".format(constraint_type))
    Gl_m = sp.vstack(Gl).tocsc()
    hl_v = np.concatenate([np.atleast_1d(np.asarray(v, dtype=float)) for v in hl])
    l = Gl_m.shape[0]
    if Gq:
        G = sp.vstack([Gl_m] + Gq).tocsc()
        h = np.concatenate([hl_v] + hq)
        nq = sum(len(v) for v in hq) // 3
    else:
        G, h, nq = Gl_m, hl_v, 0
    A = sp.vstack(Ae).tocsc() if Ae else None
    b = np.asarray(be, dtype=float) if Ae else None
    groups = np.concatenate([np.atleast_1d(np.asarray(g, dtype=int)) for g in gl])
    meta = dict(T=T, N=N, free=free, x0=x0, nfree=nfree, const=const, lb=lb, ub=ub, groups=groups,
                n_groups=T + len(rows_s))
    return P, q, G, h, l, nq, A, b, meta


def solve_mpc(objective, sessions, infra, interface, constraint_type="SOC",
              enforce_energy_equality=False, peak_limit=None, prev_peak=0, verbose=False,
              return_info=False, tol_scale=1.0):
    """Restates AdaptiveChargingOptimization.solve (aco.py:286-321): returns an
    (N, T) float64 matrix, zeros((N,1)) for no sessions, raises OracleInfeasible
    where the reference raises InfeasibilityException.  ``tol_scale`` < 1 tightens the interior-point tolerances
    (best effort: the iteration is accepted at its last iterate once the float64 Newton systems break down); the
    nearly flat objectives of the unique-optimum rate tests need it (at the default tolerances the central-path
    iterate of quick_charge + 1e-3 equal_share is still ~2e-2 A away from the optimum)."""
    if len(sessions) == 0:
        z = np.zeros((infra.num_stations, 1))
        return (z, {}) if return_info else z
    P, q, G, h, l, nq, A, b, meta = canonicalize(
        objective, sessions, infra, interface, constraint_type, enforce_energy_equality, peak_limit, prev_peak)
    # Conflicting fixed variables (e.g. min rate > infrastructure) and empty interiors
    # are decided by phase 1.
    scale = max(1.0, np.abs(h).max())
    n = len(q)
    if n == 0:
        raise OracleInfeasible("no free variables")
    lbp = violations(meta["lb"], sessions, infra, interface, constraint_type, peak_limit, enforce_energy_equality)
    if enforce_energy_equality or max(lbp["energy"], lbp["infrastructure_rel"], lbp.get("peak_rel", -1.0)) > 0:
        # the all-lower-bound schedule is not an obvious strictly feasible point: decide by phase 1
        tau, r1 = conic_ipm.phase1(G, h, l, nq, A, b, groups=meta["groups"], n_groups=meta["n_groups"],
                                   feastol=1e-8, abstol=1e-9, reltol=1e-9)
        if r1.status != "optimal" or not np.isfinite(tau) or tau > 1e-7 * scale:
            raise OracleInfeasible(f"phase-1 status {r1.status}, slack {tau:.3e}")
    res = conic_ipm.solve(P, q, G, h, l, nq, A, b, feastol=1e-9 * tol_scale, abstol=1e-8 * tol_scale, reltol=1e-9 * tol_scale,
                          verbose=verbose, max_iter=100 if tol_scale >= 1 else 200)
    if res.status != "optimal" and not (res.pres < 1e-6 and res.dres < 1e-6 and res.gap < 1e-5 * max(1, abs(res.pcost))):
        raise OracleInfeasible(f"interior-point status {res.status} (pres {res.pres:.1e}, dres {res.dres:.1e}, gap {res.gap:.1e})")
    R = meta["x0"].copy()
    R[meta["free"]] = res.x[: meta["nfree"]]
    if return_info:
        return R, dict(iters=res.iters, pcost=res.pcost + meta["const"], gap=res.gap, pres=res.pres, dres=res.dres)
    return R


def solve_lp_highs(objective, sessions, infra, interface, constraint_type="SOC",
                   enforce_energy_equality=False, peak_limit=None, prev_peak=0):
    """Independent cross-check for LP-representable instances (no quadratic terms;
    LINEAR rows, or SOC rows whose EVSEs all share one phase so that the norm is
    |sum|).  Returns (rates or None if infeasible, objective in max form)."""
    from scipy.optimize import linprog

    if constraint_type == "SOC" and has_infrastructure(infra):
        ph = np.asarray(infra.phases, dtype=float)
        A = np.asarray(infra.constraint_matrix, dtype=float)
        for j in range(A.shape[0]):
            nzp = ph[A[j] != 0]
            if len(nzp) and np.ptp(nzp) != 0:
                raise ValueError("SOC rows mix phases: not an LP")
        # |A_j x| <= lim  ->  two linear rows; build via LINEAR path on +/-A
    P, q, G, h, l, nq, A_, b_, meta = canonicalize(
        objective, sessions, infra, interface, "LINEAR" if constraint_type == "LINEAR" else "SOC",
        enforce_energy_equality, peak_limit, prev_peak)
    if P.nnz and np.abs(P.data).max() > 0:
        raise ValueError("quadratic objective: not an LP")
    Gd = G[:l]
    hd = h[:l]
    if nq:
        Gq = G[l:].tocsr()
        hq = h[l:]
        # cone k: (lim; r1; r2): single phase => r1, r2 colinear; enforce |r1|,|r2| via box on each
        # exact: r = (lim - 0, ...) ; with one phase, norm = |c| * |sum| where (cos, sin) fixed.
        rows1, rows2 = Gq[1::3], Gq[2::3]
        lim = hq[0::3]
        b1, b2 = hq[1::3], hq[2::3]
        # s1 = b1 - rows1 x, s2 = b2 - rows2 x, need hypot(s1, s2) <= lim; colinear => s = sqrt(rows1^2+rows2^2) direction
        # use combined magnitude row: for each nonzero column the pair (c0, c1) = v*(cos, sin); |v| = hypot.
        R1, R2 = rows1.tocoo(), rows2.tocoo()
        mag = sp.csr_matrix(rows1.shape)
        # sign-carrying magnitude: project onto the common phase direction of each cone
        dirc = np.zeros((rows1.shape[0], 2))
        d1 = np.asarray(abs(rows1).sum(axis=1)).ravel()
        d2 = np.asarray(abs(rows2).sum(axis=1)).ravel()
        first1 = rows1.tocsr()
        # direction from the phase of the first EVSE on the row
        ph = np.deg2rad(np.asarray(infra.phases, dtype=float))
        Am = np.asarray(infra.constraint_matrix, dtype=float)
        T = meta["T"]
        M = Am.shape[0]
        dj = np.array([[np.cos(ph[np.nonzero(Am[j])[0][0]]), np.sin(ph[np.nonzero(Am[j])[0][0]])] for j in range(M)])
        dirs = np.repeat(dj, T, axis=0)
        comb = sp.diags(dirs[:, 0]) @ rows1 + sp.diags(dirs[:, 1]) @ rows2  # = -(A_j x_t)
        bc = dirs[:, 0] * b1 + dirs[:, 1] * b2
        Gd = sp.vstack([Gd, comb, -comb]).tocsc()
        hd = np.concatenate([hd, lim + bc, lim - bc])
    res = linprog(q, A_ub=Gd, b_ub=hd, A_eq=A_, b_eq=b_, bounds=(None, None), method="highs")
    if res.status == 2:
        return None, None
    if res.status != 0:
        raise RuntimeError(res.message)
    R = meta["x0"].copy()
    R[meta["free"]] = res.x[: meta["nfree"]]
    return R, -(res.fun + meta["const"])
