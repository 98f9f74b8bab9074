"""The adapter classes beyond the plain step: preprocessing options of AdaptiveSchedulingAlgorithm
(estimate_max_rate, uninterrupted_charging: reference adacharge/adacharge.py:141-150) and the offline
algorithm (ada.py:196-294), each against the oracle on the same preprocessed sessions."""
from types import SimpleNamespace

import numpy as np
import pytest

import adacharge_b200 as ab
from adacharge_b200.adacharge import get_active_sessions
from adacharge_b200.algorithms_shim import apply_minimum_charging_rate, apply_upper_bound_estimate, enforce_pilot_limit
from adacharge_b200.generators import caltech_acn_infrastructure, config_c2, three_phase_balanced_network
from oracle import mpc

pytestmark = pytest.mark.gpu
SPEC = [("quick_charge", 1, {}), ("equal_share", 1e-3, {})]
OBJ = [ab.ObjectiveComponent(getattr(ab, n), c, k) for n, c, k in SPEC]


class _Estimator:
    def register_interface(self, interface):
        self.interface = interface

    def get_maximum_rates(self, sessions):
        return {s.session_id: 16.0 for s in sessions[::2]}


def test_estimated_max_rate_and_uninterrupted_charging(require_gpu):
    d = config_c2(17, infra=caltech_acn_infrastructure(transformer_cap=40))
    iface = ab.TestingInterface(d)
    I = iface.infrastructure_info()
    alg = ab.AdaptiveSchedulingAlgorithm(OBJ, estimate_max_rate=True, max_rate_estimator=_Estimator(), uninterrupted_charging=True)
    alg.register_interface(iface)
    sched = alg.run()
    R = np.stack([sched[s] for s in I.station_ids])
    # the same preprocessing chain on the host, then the oracle
    S = enforce_pilot_limit(iface.active_sessions(), I)
    S = apply_upper_bound_estimate(_Estimator(), S)
    S = apply_minimum_charging_rate(S, I, iface.period)
    assert any(s.min_rates[0] > 0 for s in S) and any(s.max_rates.max() <= 16 for s in S)
    T = mpc.horizon(S)
    assert R.shape == (I.num_stations, T)
    v = mpc.violations(R, S, I, iface, "SOC", None)
    assert v["lb"] <= 1e-4 and v["ub"] <= 1e-4 and v["infrastructure_rel"] <= 1e-5 and v["energy"] <= 2e-4, v
    Ro = mpc.solve_mpc(SPEC, S, I, iface, "SOC", False, None, iface.get_prev_peak())
    f, fo = (mpc.evaluate_objective(X, SPEC, I, iface, S, iface.get_prev_peak()) for X in (R, Ro))
    assert abs(f - fo) <= 1e-4 * abs(fo), (f, fo)


def test_offline_algorithm_solves_a_whole_day(require_gpu):
    """All sessions of a day in one instance, several EVs per EVSE one after the other (ada.py:234-276)."""
    rng = np.random.default_rng(3)
    infra = three_phase_balanced_network(2, 70.0)
    T, evs = 120, []
    for st in range(6):
        t = int(rng.integers(0, 10))
        k = 0
        while t + 12 < T:
            dur = int(rng.integers(12, 40))
            dep = min(t + dur, T)
            e = float(rng.uniform(2, 0.8 * (dep - t) * 32 * 208 / 1000 * 5 / 60))
            evs.append(SimpleNamespace(station_id=str(st), session_id=f"{st}-{k}", requested_energy=e, energy_delivered=0.0, arrival=t, departure=dep))
            t, k = dep + int(rng.integers(0, 6)), k + 1
    assert len(evs) > 12
    events = SimpleNamespace(queue=[(ev.arrival, SimpleNamespace(event_type="Plugin", ev=ev)) for ev in evs]
                             + [(5, SimpleNamespace(event_type="Unplug", ev=evs[0]))])
    iface = ab.TestingInterface({"active_sessions": [], "infrastructure_info": infra, "current_time": 0, "period": 5})
    alg = ab.AdaptiveChargingAlgorithmOffline(OBJ)
    with pytest.raises(ValueError):
        alg.solve()  # no interface yet
    alg.register_interface(iface)
    with pytest.raises(ValueError):
        alg.solve()  # no events yet
    alg.register_events(events)
    alg.solve()
    I = iface.infrastructure_info()
    R = np.stack([alg.internal_schedule[s] for s in I.station_ids])
    S = enforce_pilot_limit(get_active_sessions(evs, 0), I)
    assert R.shape == (6, mpc.horizon(S))
    v = mpc.violations(R, S, I, iface, "SOC", None)
    assert v["lb"] <= 1e-4 and v["ub"] <= 1e-4 and v["infrastructure_rel"] <= 1e-5 and v["energy"] <= 2e-4, v
    Ro = mpc.solve_mpc(SPEC, S, I, iface, "SOC", False, None, 0)
    f, fo = (mpc.evaluate_objective(X, SPEC, I, iface, S, 0) for X in (R, Ro))
    assert abs(f - fo) <= 1e-4 * abs(fo), (f, fo)
    iface.data["current_time"] = 30
    now = [ev for ev in evs if ev.arrival <= 30 < ev.departure]
    out = alg.schedule(now)
    assert out == {ev.station_id: [alg.internal_schedule[ev.station_id][30]] for ev in now}


def _reference_min_rate_loop(sessions, infra, override):
    """The per-session loop as acnportal writes it, one feasibility call per session."""
    from copy import deepcopy

    queue = deepcopy(sorted(sessions, key=lambda x: x.arrival))
    rates = np.zeros(len(infra.station_ids))
    flags = []
    for s in queue:
        i = infra.get_station_index(s.station_id)
        rates[i] = min(infra.min_pilot[i], override)
        ok = bool(ab.infrastructure_constraints_feasible(rates, infra))
        if not ok:
            rates[i] = 0
        flags.append(ok)
    return flags


@pytest.mark.parametrize("cap,override", [(150, float("inf")), (20, float("inf")), (12, 5), (6, 8)])
def test_min_rate_admission_matches_the_sequential_loop(require_gpu, cap, override):
    """acb_min_rate_admission (one launch) against the reference-shaped loop, from slack to heavily congested
    transformers (some sessions are refused their minimum rate)."""
    from adacharge_b200 import engine

    d = config_c2(23, infra=caltech_acn_infrastructure(transformer_cap=cap))
    iface = ab.TestingInterface(d)
    I = iface.infrastructure_info()
    S = enforce_pilot_limit(iface.active_sessions(), I)
    want = _reference_min_rate_loop(S, I, override)
    got = apply_minimum_charging_rate(S, I, override)
    queue = sorted(S, key=lambda x: x.arrival)
    assert len(got) == len(want)
    for s_in, s_out, ok in zip(queue, got, want):
        assert s_in.session_id == s_out.session_id
        i = I.get_station_index(s_in.station_id)
        if ok:
            assert s_out.min_rates[0] == max(min(I.min_pilot[i], override), s_in.min_rates[0])
            assert s_out.max_rates[0] >= s_out.min_rates[0]
        else:
            assert s_out.min_rates[0] == 0 and s_out.max_rates[0] == 0
    if cap <= 12:
        assert not all(want) and any(want)
    # batched form: the same instance three times plus a reversed offer order
    rows = np.array([[I.get_station_index(s.station_id) for s in queue]], dtype=np.int32)
    tries = np.array([[min(I.min_pilot[i], override) for i in rows[0]]])
    site = engine.get_site(I, "SOC", False, False)
    B = np.repeat(rows, 3, 0); B[2] = B[2, ::-1]
    flags = engine.min_rate_admission(site, [rows.shape[1]] * 3, B, np.repeat(tries, 3, 0))
    assert flags[0].tolist() == want and flags[1].tolist() == want
    assert flags[2].sum() > 0
