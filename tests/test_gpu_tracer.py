"""User-defined objective components given as numeric callables: the tracer recovers their packed
form, and the device solve equals the oracle's solve of the same objective written with the built-in components."""
import numpy as np
import pytest

import adacharge_b200 as ab
from adacharge_b200.generators import session_generator, three_phase_balanced_network
from oracle import mpc

pytestmark = pytest.mark.gpu


def _iface(seed, n=12, T=72):
    rng = np.random.default_rng(seed)
    infra = three_phase_balanced_network(4, 60)
    arr = rng.integers(0, T // 3, size=n)
    dep = np.minimum(arr + rng.integers(T // 4, T, size=n), T)
    dep[0] = T
    dem = rng.uniform(4, 18, size=n)
    sessions = session_generator(n, arr.tolist(), dep.tolist(), dem.tolist(), dem.tolist(), [32] * n)
    return ab.TestingInterface({"active_sessions": sessions, "infrastructure_info": infra, "current_time": 0, "period": 5,
                                "prices": (0.05 + 0.25 * rng.random(T)).tolist(), "demand_charge": 15.51, "prev_peak": 0.0})


def solar_following(rates, infrastructure, interface, solar=None, **kw):
    """-(aggregate power - solar)^2 summed over the horizon: track on-site generation."""
    u = (np.asarray(rates) * (np.asarray(infrastructure.voltages)[:, None] / 1e3)).sum(axis=0)
    return -float(((u - np.asarray(solar)[: len(u)]) ** 2).sum())


def early_energy(rates, infrastructure, interface, **kw):
    """kWh delivered, earlier periods worth more."""
    T = np.shape(rates)[1]
    e = (np.asarray(rates) * (np.asarray(infrastructure.voltages)[:, None] / 1e3)).sum(axis=0) * interface.period / 60
    return float((1.0 - 0.5 * np.arange(T) / T) @ e)


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_traced_components_solve_like_their_builtin_twins(require_gpu, seed):
    iface = _iface(seed)
    S, I = iface.active_sessions(), iface.infrastructure_info()
    T = mpc.horizon(S)
    solar = 40 * np.sin(np.linspace(0, np.pi, T)) ** 2
    obj = [ab.ObjectiveComponent(solar_following, 0.02, {"solar": solar}), ab.ObjectiveComponent(early_energy, 1.0),
           ab.ObjectiveComponent(ab.equal_share, 1e-4)]
    aco = ab.AdaptiveChargingOptimization(obj, iface)
    R = aco.solve(S, I)
    # the same objective in the oracle's vocabulary: load flattening against -solar, and a quick-charge-like linear term
    # (early_energy = sum_t c_t * period/60 * sum_i k_i r_it): evaluate both schedules with the numeric callables themselves
    k = np.asarray(I.voltages) / 1e3
    c = (1.0 - 0.5 * np.arange(T) / T) * iface.period / 60
    spec = [("load_flattening", 0.02, {"external_signal": -solar}), ("equal_share", 1e-4, {}),
            ("linear", 1.0, {"weights": k[:, None] * c[None, :]})]

    def value(X):
        return 0.02 * solar_following(X, I, iface, solar=solar) + early_energy(X, I, iface) - 1e-4 * float((X ** 2).sum())

    Ro = mpc.solve_mpc(spec, S, I, iface, "SOC", False, None, 0.0)
    f, fo = value(R), value(Ro)
    # the oracle's own evaluation of its spec agrees with the callables
    assert mpc.evaluate_objective(Ro, spec, I, iface, S) == pytest.approx(fo, rel=1e-9)
    assert abs(f - fo) <= 1e-4 * abs(fo) + 1e-7, (f, fo, aco.last_info)
    v = mpc.violations(R, S, I, iface)
    assert v["infrastructure_rel"] <= 1e-5 and v["lb"] <= 0 and v["ub"] <= 0 and v["energy"] <= 1e-4, v


@pytest.mark.parametrize("scale", [0.3, 0.75, 1.5], ids=["optimum_above_both", "optimum_between", "optimum_below_both"])
def test_peak_terms_with_different_baselines(require_gpu, scale):
    """Two demand-charge components with their own baseline_peak (aco.py:387-400): the sum of two epigraphs with
    different kinks.  The device objective holds one linear piece of it at a time; solve() walks the pieces.  The
    baselines are set relative to the peak of the plain (single baseline 0) optimum so that the three cases are hit."""
    iface = _iface(5, n=12, T=60)
    S, I = iface.active_sessions(), iface.infrastructure_info()
    base = [("tou_energy_cost", 1, {}), ("total_energy", 0.4, {}), ("equal_share", 1e-5, {})]
    R0 = mpc.solve_mpc(base + [("demand_charge", 0.05, {})], S, I, iface, "SOC", False, None, 0.0)
    k = np.asarray(I.voltages) / 1e3
    m0 = (k @ R0).max()
    p1, p2 = 0.8 * scale * m0, 1.3 * scale * m0
    spec = base + [("demand_charge", 0.03, {"baseline_peak": p1}), ("demand_charge", 0.04, {"baseline_peak": p2})]
    obj = [ab.ObjectiveComponent(getattr(ab, n), c, kw) for n, c, kw in spec]
    aco = ab.AdaptiveChargingOptimization(obj, iface)
    R = aco.solve(S, I)
    Ro = mpc.solve_mpc(spec, S, I, iface, "SOC", False, None, 0.0)
    f, fo = (mpc.evaluate_objective(X, spec, I, iface, S, 0.0) for X in (R, Ro))
    assert abs(f - fo) <= 1e-4 * abs(fo) + 1e-7, (f, fo, (k @ R).max(), (k @ Ro).max(), p1, p2, aco.last_info)
    v = mpc.violations(R, S, I, iface)
    assert v["infrastructure_rel"] <= 1e-5 and v["lb"] <= 0 and v["ub"] <= 0 and v["energy"] <= 1e-4, v
