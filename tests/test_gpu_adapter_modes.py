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


def _numpy_feasible(rates, infra):
    """infrastructure_constraints_feasible in plain numpy (reference adacharge/utils.py:5-12): no device code."""
    ph = np.deg2rad(np.asarray(infra.phases, dtype=float))
    for j, v in enumerate(np.asarray(infra.constraint_matrix, dtype=float)):
        a = np.stack([v * np.cos(ph), v * np.sin(ph)])
        if not np.all(np.linalg.norm(a @ rates, axis=0) <= infra.constraint_limits[j] + 1e-7):
            return False
    return True


def _reference_min_rate_loop(sessions, infra, override):
    """The per-session loop as acnportal writes it, one feasibility call per session, in pure numpy."""
    from copy import deepcopy

    queue = deepcopy(sorted(sessions, key=lambda x: x.arrival))
    rates = np.zeros(len(infra.station_ids))
    flags = []
    for s in queue:
        i = infra.get_station_index(s.station_id)
        rates[i] = min(infra.min_pilot[i], override)
        ok = _numpy_feasible(rates, infra)
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


def test_batched_device_preprocessing_matches_the_host_helpers(require_gpu):
    """acb_preprocess_sessions (enforce_pilot_limit + apply_upper_bound_estimate on the raw [B, S] tables, in place)
    against the per-instance host helpers of the adapter (reference adacharge/adacharge.py:141-146)."""
    import ctypes as C

    import torch

    from adacharge_b200 import _cabi, engine
    from adacharge_b200.batched import BatchedAdaptiveCharging, sessions_to_arrays
    from adacharge_b200.algorithms_shim import apply_upper_bound_estimate

    infra = caltech_acn_infrastructure()
    infra["max_pilot"] = np.where(np.arange(54) % 3 == 0, 16.0, 32.0)  # some EVSEs with a smaller pilot limit
    rng = np.random.default_rng(3)
    ifaces = [ab.TestingInterface(config_c2(90 + i, infra=infra)) for i in range(6)]
    I = ifaces[0].infrastructure_info()
    lists = [f.active_sessions() for f in ifaces]
    for ss in lists:  # a few sessions with a minimum rate, so that the reconcile rule matters
        for s in ss[::5]:
            s.min_rates = np.full(s.remaining_time, 8.0)
    sess = sessions_to_arrays(lists, I, S_max=54)
    upper = rng.uniform(4.0, 40.0, size=sess["max_rate"].shape)

    class Est:
        def __init__(self, ss, row): self.m = {s.session_id: row[j] for j, s in enumerate(ss)}
        def get_maximum_rates(self, sessions): return self.m

    bac = BatchedAdaptiveCharging([ab.ObjectiveComponent(ab.quick_charge)], I, 5, batch=6, max_sessions=54, horizon=288, chunks=1,
                                  enforce_pilot_limit=True, estimate_max_rate=True)
    ch = bac.chunks[0]
    bac.upload_raw(sess, upper_bound=upper)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    _cabi.check(_cabi.lib().acb_preprocess_sessions(bac.site.handle, C.byref(ch.sessions), 1, engine._ptr(ch.dev_raw["upper_bound"]), st), "pre")
    got = ch.dev_raw["max_rate"].cpu().numpy()
    for b, ss in enumerate(lists):
        want = apply_upper_bound_estimate(Est(ss, upper[b]), enforce_pilot_limit(ss, I))
        for j, s in enumerate(want):
            assert np.all(s.max_rates == got[b, j]), (b, j, s.max_rates[:2], got[b, j])
    assert (got[sess["station"] < 0] == 0).all()
