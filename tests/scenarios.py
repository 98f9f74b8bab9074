"""The scenarios of the reference's solver tests
(reference adacharge/tests/test_adaptive_charging_optimization.py), as data, shared by
the oracle tests (CPU) and the CUDA parity tests (GPU)."""
import numpy as np

from adacharge_b200.generators import session_generator, single_phase_single_constraint, three_phase_balanced_network

PERIOD, MAX_RATE, ENERGY, HORIZON = 5, 32, 3.3, 12


def _tiny(arrivals=(0, 0), departures=(HORIZON, HORIZON), station_ids=None, min_rates=None, limit=64):
    s = session_generator(2, list(arrivals), list(departures), [ENERGY] * 2, [ENERGY] * 2, [MAX_RATE] * 2,
                          min_rates=min_rates, station_ids=station_ids)
    return s, single_phase_single_constraint(2, limit)


def _large(three_phase):
    n, T = 54, 144
    s = session_generator(n, [0] * n, [T] * n, [10] * n, [10] * n, [MAX_RATE] * n)
    infra = three_phase_balanced_network(n // 3, 32 * n / 3) if three_phase else single_phase_single_constraint(n, 32 * n / 3)
    return s, infra


QC = [("quick_charge", 1, {})]
# name -> dict(sessions, infra, objective, constraint_type, equality, peak_limit, current_time, prices, checks)
SCENARIOS = {
    "tiny_feasible": dict(data=_tiny(), objective=QC),                                                   # t_aco.py:87-100
    "tiny_energy_equality": dict(data=_tiny(), objective=QC, equality=True),                             # :103-116
    "tiny_delayed_start": dict(data=_tiny(arrivals=(0, 4), departures=(HORIZON, HORIZON + 4)), objective=QC),  # :178-191
    "tiny_same_evse": dict(data=_tiny(arrivals=(0, 12), departures=(HORIZON, HORIZON + 12), station_ids=["0", "0"]), objective=QC),  # :194-208
    "tiny_min_charge": dict(data=_tiny(min_rates=[6, 6]), objective=QC, min_rate=6),                     # :211-229
    "tiny_peak_scalar": dict(data=_tiny(), objective=QC, peak_limit=32),                                 # :232-257
    "tiny_peak_vector": dict(data=_tiny(), objective=QC, peak_limit=np.array([40] * 6 + [24] * 6)),      # :260-282
    "large_single_linear": dict(data=_large(False), objective=QC, constraint_type="LINEAR", max_energy=10),  # :286-313
    "large_single_soc": dict(data=_large(False), objective=QC, max_energy=10),                           # :316-343
    "large_three_soc": dict(data=_large(True), objective=QC, max_energy=10),                             # :374-403
    "large_three_equal_share": dict(data=_large(True), objective=QC + [("equal_share", 1e-12, {})], max_energy=10),  # :406-435
    "large_three_linear": dict(data=_large(True), objective=QC, constraint_type="LINEAR", max_energy=10),  # :438-466
    "tou_tiny": dict(data=_tiny(), objective=[("tou_energy_cost", 1, {})], equality=True,
                     prices=[0.3] * 6 + [0.1] * 6, no_charge_cols=6),                                     # :469-507
    "tou_tiny_t4": dict(data=_tiny(), objective=[("tou_energy_cost", 1, {})], equality=True, current_time=4,
                        prices=[0.3] * 2 + [0.1] * 6, no_charge_cols=2, positive_after=True),             # :510-545
}
INFEASIBLE = {
    "infeasible_max_rate": dict(data=_tiny(departures=(12, 4)), objective=QC, equality=True),            # :119-145
    "infeasible_infrastructure": dict(data=_tiny(limit=30), objective=QC, equality=True),                # :148-175
}


def make_interface(sc):
    from adacharge_b200.interface import TestingInterface

    sessions, infra = sc["data"]
    ct = sc.get("current_time", 0)
    d = {"active_sessions": sessions, "infrastructure_info": infra, "current_time": ct, "period": PERIOD}
    d.update(sc.get("iface_extra", {}))
    iface = TestingInterface(d)
    if "prices" in sc:
        prices = np.array(sc["prices"], dtype=float)
        iface.get_prices = lambda length, start=None: prices[:length]  # the reference mocks get_prices the same way
    return iface


def check_properties(rates, sc, iface, tol_rate=1e-3, tol_line=1e-3):
    """The four inherited checks of the reference base class (t_aco.py:50-83) plus the
    scenario-specific ones."""
    S, I = iface.active_sessions(), iface.infrastructure_info()
    assert (rates <= MAX_RATE + tol_rate).all()
    expected = np.zeros(rates.shape[0])
    delivered = np.zeros(rates.shape[0])
    for s in S:
        i = I.station_ids.index(s.station_id)
        expected[i] = s.remaining_demand
        delivered[i] = rates[i, s.arrival_offset : s.arrival_offset + s.remaining_time].sum() * I.voltages[i] * PERIOD / 1e3 / 60
    assert np.allclose(delivered, expected, atol=1e-4, rtol=1e-4), (delivered, expected)
    unplugged = np.ones(rates.shape, dtype=bool)
    for s in S:
        i = I.station_ids.index(s.station_id)
        unplugged[i, s.arrival_offset : s.arrival_offset + s.remaining_time] = False
    assert np.allclose(rates[unplugged], 0)
    ph = np.deg2rad(I.phases)
    for j, v in enumerate(I.constraint_matrix):
        a = np.stack([v * np.cos(ph), v * np.sin(ph)])
        assert np.all(np.linalg.norm(a @ rates, axis=0) <= I.constraint_limits[j] + tol_line)
    if "min_rate" in sc:
        assert (rates >= sc["min_rate"] - 1e-7).all()
    if "peak_limit" in sc:
        assert (rates.sum(axis=0) <= np.asarray(sc["peak_limit"]) + 1e-7 * 1).all() or \
               (rates.sum(axis=0) <= np.asarray(sc["peak_limit"]) * (1 + 1e-5)).all()
    if "no_charge_cols" in sc:
        assert np.allclose(rates[:, : sc["no_charge_cols"]], 0, atol=1e-3)
        if sc.get("positive_after"):
            assert np.all(rates[:, sc["no_charge_cols"] :] > 1e-4)


def random_scenario(seed):
    """Seeded small instance mixing everything the path supports: single-/three-phase networks, SOC or
    LINEAR rows, staggered windows, a second session on an EVSE, minimum rates, energy equality, scalar
    or vector peak limits and random objective mixes (incl. the aggregate-quadratic and peak terms)."""
    rng = np.random.default_rng(1000 + seed)
    three = bool(rng.integers(0, 2))
    if three:
        per = int(rng.integers(1, 4))
        n = 3 * per
    else:
        n = int(rng.integers(2, 10))
    T = int(rng.integers(8, 40))
    equality = rng.random() < 0.25
    with_min = (not equality) and rng.random() < 0.2
    tight = rng.uniform(0.7, 0.95) if (equality or with_min) else rng.uniform(0.3, 0.9)
    limit = tight * 32 * n / (3 if three else 1) * (1.5 if three else 1.0)
    infra = three_phase_balanced_network(n // 3, limit) if three else single_phase_single_constraint(n, limit)
    k = int(rng.integers(1, n + 1))
    stations = [str(i) for i in rng.permutation(n)[:k]]
    arr = rng.integers(0, max(1, T // 2), k)
    dep = np.minimum(arr + rng.integers(3, T, k), T)
    dep[rng.integers(0, k)] = T
    deliverable = (dep - arr) * 32 * 208 / 1000 * PERIOD / 60
    frac = rng.uniform(0.05, 0.3, k) if equality else rng.uniform(0.1, 0.9, k)
    energy = frac * deliverable
    arr, dep, energy = list(map(int, arr)), list(map(int, dep)), list(map(float, energy))
    if rng.random() < 0.3 and dep[0] + 3 <= T:  # a second session on the first EVSE, after the first one left
        stations.append(stations[0]); arr.append(dep[0]); dep.append(T); energy.append(float(0.2 * (T - dep[0]) * 32 * 208 / 1000 * PERIOD / 60))
    m = len(stations)
    mins = [6.0] * m if with_min else None
    if with_min:
        energy = [max(e, 1.05 * 6 * (d - a) * 208 / 1000 * PERIOD / 60) for e, a, d in zip(energy, arr, dep)]
    sessions = session_generator(m, arr, dep, energy, energy, [MAX_RATE] * m, min_rates=mins, station_ids=stations)
    obj = []
    pool = rng.permutation(5)[: int(rng.integers(1, 4))]
    for p in pool:
        if p == 0:
            obj.append(("quick_charge", 1.0, {}))
        elif p == 1:
            obj.append(("equal_share", float(rng.choice([1e-3, 1e-2, 0.05])), {}))
        elif p == 2:
            obj += [("tou_energy_cost", 1.0, {}), ("total_energy", 0.3, {})]
        elif p == 3:
            obj.append(("demand_charge", 1 / 30, {}))
            if not any(o[0] in ("quick_charge", "total_energy") for o in obj):
                obj.append(("total_energy", 0.3, {}))
        else:
            obj.append(("load_flattening", 1e-3, {"external_signal": (20 + 10 * np.sin(np.arange(T) / 5.0)).tolist()}))
            if not any(o[0] in ("quick_charge", "total_energy") for o in obj):
                obj.append(("quick_charge", 1.0, {}))
    sc = dict(data=(sessions, infra), objective=obj, constraint_type="LINEAR" if rng.random() < 0.4 else "SOC", equality=equality,
              iface_extra=dict(prices=(0.05 + 0.25 * rng.random(T + 4)).tolist(), demand_charge=15.51,
                               prev_peak=float(rng.choice([0.0, 0.0, 20.0, 60.0]))))
    if not (equality or with_min) and rng.random() < 0.3:
        cap = 0.6 * 32 * m
        sc["peak_limit"] = float(cap) if rng.random() < 0.5 else np.linspace(cap, 0.8 * cap, T)
    return sc
