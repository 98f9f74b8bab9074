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
