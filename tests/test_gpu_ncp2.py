"""non_completion_penalty with norm = 2 (build-defined component, SURVEY.md 8(a) A14: absent from the reference): the
device solver treats it as 'soft' energy rows -- a per-session quadratic in the planned energy, whose prox keeps the
one-multiplier structure of the energy-row projection -- on both kernels.  Checked against the oracle's epigraph form."""
import numpy as np
import pytest

import adacharge_b200 as ab
from adacharge_b200.generators import session_generator, three_phase_balanced_network, single_phase_single_constraint
from oracle import mpc

pytestmark = pytest.mark.gpu


def _case(kind, seed):
    rng = np.random.default_rng(seed)
    if kind == "three_phase":
        n, T = 9, 60
        infra = three_phase_balanced_network(3, 40)  # tight lines: not every request can be met
    else:
        n, T = 6, 48
        infra = single_phase_single_constraint(n, 50)
    arr = rng.integers(0, T // 3, size=n)
    dep = np.minimum(arr + rng.integers(T // 4, T, size=n), T)
    dep[0] = T
    dem = rng.uniform(8, 30, size=n)  # kWh: more than the windows can deliver for several sessions
    sessions = session_generator(n, arr.tolist(), dep.tolist(), dem.tolist(), dem.tolist(), [32] * n)
    iface = ab.TestingInterface({"active_sessions": sessions, "infrastructure_info": infra, "current_time": 0, "period": 5,
                                 "prices": (0.05 + 0.25 * rng.random(T)).tolist(), "demand_charge": 15.51, "prev_peak": 0.0})
    return iface


@pytest.mark.parametrize("path", [1, 2], ids=["on_chip", "general"])
@pytest.mark.parametrize("kind,seed,spec", [
    ("three_phase", 1, [("tou_energy_cost", 1, {}), ("non_completion_penalty", 0.05, {"norm": 2})]),
    ("three_phase", 2, [("tou_energy_cost", 1, {}), ("non_completion_penalty", 0.5, {"norm": 2}), ("demand_charge", 0.02, {})]),
    ("single_phase", 3, [("quick_charge", 0.01, {}), ("non_completion_penalty", 0.2, {"norm": 2}), ("equal_share", 1e-3, {})]),
    ("single_phase", 4, [("tou_energy_cost", 1, {}), ("non_completion_penalty", 0.02, {"norm": 2}), ("non_completion_penalty", 0.1, {"norm": 1})]),
])
def test_quadratic_non_completion_penalty_matches_oracle(require_gpu, path, kind, seed, spec):
    iface = _case(kind, seed)
    S, I = iface.active_sessions(), iface.infrastructure_info()
    obj = [ab.ObjectiveComponent(getattr(ab, n), c, k) for n, c, k in spec]
    aco = ab.AdaptiveChargingOptimization(obj, iface, solver_options=dict(path=path))
    R = aco.solve(S, I, prev_peak=iface.get_prev_peak())
    Ro = mpc.solve_mpc(spec, S, I, iface, "SOC", False, None, iface.get_prev_peak())
    f, fo = (mpc.evaluate_objective(X, spec, I, iface, S, iface.get_prev_peak()) for X in (R, Ro))
    assert abs(f - fo) <= 1e-4 * abs(fo), (f, fo, aco.last_info)
    v = mpc.violations(R, S, I, iface)
    assert v["infrastructure_rel"] <= 1e-5 and v["lb"] <= 0 and v["ub"] <= 0 and v["energy"] <= 1e-4, v
    # the penalty is doing something: some session is left short of its request
    w = np.asarray(I.voltages) * 5 / 1e3 / 60
    short = [s.remaining_demand - w[I.get_station_index(s.station_id)] * R[I.get_station_index(s.station_id), s.arrival_offset:s.arrival_offset + s.remaining_time].sum() for s in S]
    assert max(short) > 1e-2
