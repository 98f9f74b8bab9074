"""Pins the MPC oracle: every scenario of the reference's solver tests, analytic unique
optima, and HiGHS on LP-representable cases (CPU only)."""
import numpy as np
import pytest

from oracle import mpc
from tests.scenarios import SCENARIOS, INFEASIBLE, make_interface, check_properties

FAST = [k for k in SCENARIOS if not k.startswith("large")]
LARGE = [k for k in SCENARIOS if k.startswith("large")]


def _solve(sc):
    iface = make_interface(sc)
    S, I = iface.active_sessions(), iface.infrastructure_info()
    R = mpc.solve_mpc(sc["objective"], S, I, iface, sc.get("constraint_type", "SOC"), sc.get("equality", False),
                      sc.get("peak_limit"), 0)
    return R, iface, S, I


@pytest.mark.parametrize("name", FAST + LARGE)
def test_reference_scenarios_hold_for_oracle(name):
    sc = SCENARIOS[name]
    R, iface, S, I = _solve(sc)
    if "peak_limit" in sc:
        assert (R.sum(axis=0) <= np.asarray(sc["peak_limit"]) + 1e-7).all()  # t_aco.py:256-257
    check_properties(R, sc, iface)


@pytest.mark.parametrize("name", list(INFEASIBLE))
def test_infeasible_scenarios_raise(name):
    sc = INFEASIBLE[name]
    with pytest.raises(mpc.OracleInfeasible):
        _solve(sc)


def test_kat1_analytic_optimum():
    R, iface, S, I = _solve(SCENARIOS["tiny_feasible"])
    w = 208 * 5 / 1e3 / 60
    need = 3.3 / w  # A*periods
    row = np.array([32.0] * 5 + [need - 160] + [0.0] * 6)
    assert np.allclose(R, np.stack([row, row]), atol=1e-5)
    assert abs(mpc.evaluate_objective(R, SCENARIOS["tiny_feasible"]["objective"], I, iface) - 302.1153846) < 1e-5


@pytest.mark.parametrize("name", ["tiny_feasible", "tiny_delayed_start", "tiny_min_charge", "tiny_peak_scalar",
                                  "tiny_peak_vector", "large_single_linear", "large_single_soc", "large_three_linear"])
def test_objective_matches_highs(name):
    sc = SCENARIOS[name]
    R, iface, S, I = _solve(sc)
    Rh, fh = mpc.solve_lp_highs(sc["objective"], S, I, iface, sc.get("constraint_type", "SOC"), sc.get("equality", False),
                                sc.get("peak_limit"), 0)
    f = mpc.evaluate_objective(R, sc["objective"], I, iface)
    assert abs(f - fh) <= 1e-6 * max(1.0, abs(fh))


def test_empty_sessions_returns_zero_column():
    iface = make_interface(SCENARIOS["tiny_feasible"])
    I = iface.infrastructure_info()
    assert mpc.solve_mpc([("quick_charge", 1, {})], [], I, iface).shape == (2, 1)


def test_objective_components_against_direct_formulas():
    from adacharge_b200.generators import config_c2
    from adacharge_b200.interface import TestingInterface

    d = config_c2(5)
    iface = TestingInterface(d)
    S, I = iface.active_sessions(), iface.infrastructure_info()
    T = mpc.horizon(S)
    rng = np.random.default_rng(0)
    R = rng.uniform(0, 32, (54, T))
    k = np.asarray(I.voltages) / 1e3
    u = k @ R
    ext = rng.uniform(0, 50, T)
    obj = [("quick_charge", 2.0, {}), ("equal_share", 0.1, {}), ("tou_energy_cost", 1.5, {}), ("total_energy", 3.0, {}),
           ("demand_charge", 0.7, {}), ("load_flattening", 0.2, {"external_signal": ext}), ("non_completion_penalty", 0.4, {})]
    c = np.array([(T - t) / T for t in range(T)])
    prices = iface.get_prices(T)
    prev = iface.get_prev_peak() * 208 / 1000
    unmet = sum(abs(s.remaining_demand - 208 * 5 / 1e3 / 60 * R[I.get_station_index(s.station_id), s.arrival_offset:s.arrival_offset + s.remaining_time].sum()) for s in S)
    direct = (2.0 * (c @ R.sum(axis=0)) - 0.1 * (R ** 2).sum() - 1.5 * (prices @ (u * 5 / 60)) + 3.0 * (u * 5 / 60).sum()
              - 0.7 * 15.51 * max(u.max(), prev) - 0.2 * ((u + ext) ** 2).sum() - 0.4 * unmet)
    assert abs(mpc.evaluate_objective(R, obj, I, iface, S, iface.get_prev_peak()) - direct) < 1e-8 * abs(direct)


def _lp_random_seeds(limit=12):
    """Seeds of tests/scenarios.py::random_scenario that HiGHS can take (no quadratic term; LINEAR rows or single-phase SOC)."""
    from tests.scenarios import random_scenario

    out = []
    for seed in range(80):
        sc = random_scenario(seed)
        if any(o[0] in ("equal_share", "load_flattening") for o in sc["objective"]):
            continue
        if sc["constraint_type"] == "SOC" and len(set(np.asarray(sc["data"][1]["phases"]).tolist())) > 1:
            continue
        out.append(seed)
        if len(out) == limit:
            break
    return out


@pytest.mark.parametrize("seed", _lp_random_seeds())
def test_random_lp_instances_match_highs(seed):
    """The interior-point oracle against HiGHS on random instances with staggered windows, repeated EVSEs,
    minimum rates, equality rows, peak limits, TOU prices and the demand-charge epigraph."""
    from tests.scenarios import random_scenario

    sc = random_scenario(seed)
    iface = make_interface(sc)
    S, I = iface.active_sessions(), iface.infrastructure_info()
    args = (sc["objective"], S, I, iface, sc["constraint_type"], sc.get("equality", False), sc.get("peak_limit"), iface.get_prev_peak())
    R = mpc.solve_mpc(*args)
    Rh, fh = mpc.solve_lp_highs(*args)
    f = mpc.evaluate_objective(R, sc["objective"], I, iface, S, iface.get_prev_peak())
    assert abs(f - fh) <= 1e-6 * max(1.0, abs(fh)), (f, fh)


def _soc_random_seeds(limit=14):
    from tests.scenarios import random_scenario

    out = []
    for seed in range(120):
        sc = random_scenario(seed)
        if sc["constraint_type"] != "SOC" or len(set(np.asarray(sc["data"][1]["phases"]).tolist())) < 2:
            continue
        iface = make_interface(sc)
        if iface.infrastructure_info().num_stations * mpc.horizon(iface.active_sessions()) > 130:
            continue
        out.append(seed)
        if len(out) == limit:
            break
    return out


def _independent_soc_solve(sc):
    """The scenario solved without the oracle's canonicalisation or interior-point method: scipy SLSQP on the
    reference formulation written out directly (norm-squared line-current constraints with analytic Jacobians, the
    objective as the quadratic it is for these components)."""
    from scipy.optimize import minimize

    iface = make_interface(sc); S, I = iface.active_sessions(), iface.infrastructure_info()
    T = mpc.horizon(S); N = I.num_stations; n = N * T
    lb, ub = mpc.bounds(S, I.station_ids, T); ub = np.maximum(ub, lb)
    pp = iface.get_prev_peak()
    obj = [o for o in sc["objective"] if o[0] not in ("demand_charge", "load_flattening")]
    flat = [(o[1], np.asarray(o[2].get("external_signal", np.zeros(T)), float)[:T]) for o in sc["objective"] if o[0] == "load_flattening"]
    # demand_charge = -dc * max(max_t u_t, prev_peak kW): an epigraph variable tau appended to x (aco.py:387-400)
    w_peak = sum(o[1] for o in sc["objective"] if o[0] == "demand_charge") * iface.get_demand_charge()
    k = np.asarray(I.voltages, float) / 1000.0
    p0 = pp * I.voltages[0] / 1000.0
    F = lambda x: -mpc.evaluate_objective(x.reshape(N, T), obj, I, iface, S, pp) if obj else 0.0
    # the admitted objectives are linear + diagonal quadratic: recover the coefficients from 2n evaluations around a point inside the box
    xm = (0.5 * (lb + ub)).ravel(); f0 = F(xm); a = np.zeros(n); q = np.zeros(n)
    for i in range(n):
        e = np.zeros(n); e[i] = 1.0
        fp, fm = F(xm + e), F(xm - e)
        q[i] = 0.5 * (fp + fm) - f0; a[i] = 0.5 * (fp - fm)
    kk = np.asarray(I.voltages, float) / 1000.0

    def f(x):  # linear + diagonal quadratic part, plus load_flattening written out: coef * sum_t (u_t + ext_t)^2  (aco.py:403-408)
        val = f0 + a @ (x - xm) + q @ (x - xm) ** 2
        for coef, ext in flat:
            val += coef * ((kk @ x.reshape(N, T) + ext) ** 2).sum()
        return val

    def g(x):
        grad = a + 2 * q * (x - xm)
        for coef, ext in flat:
            grad = grad + (2 * coef * np.outer(kk, kk @ x.reshape(N, T) + ext)).ravel()
        return grad
    cons = []
    for (i, s0, s1, w, e) in mpc.session_rows(S, I, iface.period):
        row = np.zeros((N, T)); row[i, s0:s1] = w; row = row.ravel()
        if sc.get("equality"): cons.append(dict(type="eq", fun=lambda x, r=row, e=e: r @ x - e, jac=lambda x, r=row: r))
        else: cons.append(dict(type="ineq", fun=lambda x, r=row, e=e: e - r @ x, jac=lambda x, r=row: -r))
    rows = mpc.soc_rows(I); lim = np.asarray(I.constraint_limits, float); M = rows.shape[0]
    def soc(x):
        cur = np.einsum("mkn,nt->mkt", rows, x.reshape(N, T))
        return (lim[:, None] ** 2 - (cur ** 2).sum(axis=1)).ravel()
    def soc_jac(x):
        cur = np.einsum("mkn,nt->mkt", rows, x.reshape(N, T))
        J = np.zeros((M, T, N, T))
        for t in range(T):
            J[:, t, :, t] = -2 * np.einsum("mk,mkn->mn", cur[:, :, t], rows)
        return J.reshape(M * T, n)
    cons.append(dict(type="ineq", fun=soc, jac=soc_jac))
    if sc.get("peak_limit") is not None:
        pl = np.broadcast_to(np.asarray(sc["peak_limit"], float), (T,)).copy()
        A = np.zeros((T, n))
        for t in range(T): A[t, t::T] = 1
        cons.append(dict(type="ineq", fun=lambda x: pl - A @ x, jac=lambda x: -A))
    bnds = list(zip(lb.ravel(), ub.ravel()))
    if w_peak > 0:
        # variables (x, tau): every constraint so far ignores tau; add u_t <= tau, tau >= p0 and the cost w_peak * tau
        pad = lambda c: dict(type=c["type"], fun=lambda z, c=c: c["fun"](z[:n]), jac=lambda z, c=c: np.hstack([np.atleast_2d(c["jac"](z[:n])), np.zeros((np.atleast_2d(c["jac"](z[:n])).shape[0], 1))]))  # noqa: E731
        cons = [pad(c) for c in cons]
        U = np.zeros((T, n))
        for t in range(T):
            U[t, t::T] = k
        cons.append(dict(type="ineq", fun=lambda z: z[n] - U @ z[:n], jac=lambda z: np.hstack([-U, np.ones((T, 1))])))
        f_, g_ = f, g
        f = lambda z: f_(z[:n]) + w_peak * z[n]  # noqa: E731
        g = lambda z: np.r_[g_(z[:n]), w_peak]  # noqa: E731
        bnds = bnds + [(p0, None)]
        z0 = np.r_[lb.ravel(), max(p0, float((U @ lb.ravel()).max()))]
    else:
        z0 = lb.ravel().copy()
    res = minimize(f, z0, jac=g, method="SLSQP", bounds=bnds, constraints=cons, options=dict(maxiter=3000, ftol=1e-14))
    assert abs(F(res.x[:n]) - (f0 + a @ (res.x[:n] - xm) + q @ (res.x[:n] - xm) ** 2)) <= 1e-7 * max(1, abs(f0))  # the quadratic model is the objective
    return -res.fun, res, (iface, S, I)



@pytest.mark.parametrize("seed", _soc_random_seeds())
def test_mixed_phase_soc_instances_match_slsqp(seed):
    """True second-order-cone cases (three-phase rows mixing phase angles), where HiGHS cannot help: the oracle
    against an independent SLSQP solve of the directly written problem."""
    from tests.scenarios import random_scenario

    sc = random_scenario(seed)
    fs, res, (iface, S, I) = _independent_soc_solve(sc)
    R = mpc.solve_mpc(sc["objective"], S, I, iface, "SOC", sc.get("equality", False), sc.get("peak_limit"), iface.get_prev_peak())
    fo = mpc.evaluate_objective(R, sc["objective"], I, iface, S, iface.get_prev_peak())
    assert abs(fs - fo) <= 1e-6 * max(1.0, abs(fo)), (fs, fo, res.status)


# ---- the real reference, where it can run (cvxpy importable): pins the oracle to reference OUTPUTS --------------------------
from oracle import reference_cvxpy  # noqa: E402


def test_reference_hook_reports_availability_consistently():
    """available() must agree with what the image has: without cvxpy the hook stays off and nothing else is touched."""
    try:
        import cvxpy  # noqa: F401
        have = True
    except Exception:
        have = False
    assert reference_cvxpy.available() == (have and reference_cvxpy._ref_dir() is not None)


@pytest.mark.skipif(not reference_cvxpy.available(), reason="cvxpy (the reference's solver front end) is not installed in this image")
@pytest.mark.parametrize("name", FAST + LARGE)
def test_oracle_matches_the_reference_solver(name):
    """Objective of the oracle's schedule within 1e-6 |f*| of the reference's own solve (aco.py:286-321) on the same scenario."""
    sc = SCENARIOS[name]
    R, iface, S, I = _solve(sc)
    Rref = reference_cvxpy.solve_reference(sc["objective"], S, I, iface, sc.get("constraint_type", "SOC"), sc.get("equality", False),
                                           sc.get("peak_limit"), 0)
    f = mpc.evaluate_objective(R, sc["objective"], I, iface)
    fref = mpc.evaluate_objective(np.asarray(Rref), sc["objective"], I, iface)
    assert abs(f - fref) <= 1e-6 * max(1.0, abs(fref)), (f, fref)
