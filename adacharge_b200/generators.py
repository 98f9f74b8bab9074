"""Seeded synthetic sites and sessions for tests and benchmarks.

The first three functions restate the acnportal test-case generators the reference
tests import (``from acnportal.algorithms.tests.generate_test_cases import *`` —
reference tests/test_adaptive_charging_optimization.py:4, tests/test_postprocessing.py:7-12);
[acnportal, recalled], checked against the reference's known answers in
tests/test_oracle_postprocessing.py.  ``caltech_acn_infrastructure`` restates the
54-EVSE / 8-constraint Caltech site of acnportal ``sites.caltech_acn`` [recalled].
The ``config_c*`` functions are the BASELINE.json workloads.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import numpy as np


# --------------------------------------------------------------------------- fixtures
def session_generator(
    num_sessions,
    arrivals,
    departures,
    requested_energy,
    remaining_energy,
    max_rates,
    min_rates=None,
    station_ids=None,
    estimated_departures=None,
) -> List[Dict]:
    sessions = []
    for i in range(num_sessions):
        sessions.append(
            {
                "station_id": station_ids[i] if station_ids is not None else f"{i}",
                "session_id": f"{i}",
                "requested_energy": requested_energy[i],
                "energy_delivered": requested_energy[i] - remaining_energy[i],
                "arrival": arrivals[i],
                "departure": departures[i],
                "estimated_departure": (
                    estimated_departures[i]
                    if estimated_departures is not None
                    else departures[i]
                ),
                "min_rates": min_rates[i] if min_rates is not None else 0,
                "max_rates": max_rates[i],
            }
        )
    return sessions


def _default_pilots(n, min_pilot, max_pilot):
    return [np.array([0] + list(range(min_pilot, max_pilot + 1))) for _ in range(n)]


def single_phase_single_constraint(
    num_evses, limit, max_pilot=32, min_pilot=8, allowable_pilots=None, is_continuous=None
) -> Dict:
    if allowable_pilots is None:
        allowable_pilots = _default_pilots(num_evses, min_pilot, max_pilot)
    if is_continuous is None:
        is_continuous = np.ones(num_evses, dtype=bool)
    return {
        "constraint_matrix": np.ones((1, num_evses)),
        "constraint_limits": np.array([limit]),
        "phases": np.zeros(num_evses),
        "voltages": np.repeat(208, num_evses),
        "constraint_ids": ["all"],
        "station_ids": [f"{i}" for i in range(num_evses)],
        "max_pilot": np.repeat(max_pilot, num_evses),
        "min_pilot": np.repeat(min_pilot, num_evses),
        "allowable_pilots": allowable_pilots,
        "is_continuous": is_continuous,
    }


def three_phase_balanced_network(
    evses_per_phase, limit, max_pilot=32, min_pilot=8, allowable_pilots=None, is_continuous=None
) -> Dict:
    n = 3 * evses_per_phase
    if allowable_pilots is None:
        allowable_pilots = _default_pilots(n, min_pilot, max_pilot)
    if is_continuous is None:
        is_continuous = np.ones(n, dtype=bool)
    return {
        "constraint_matrix": np.array(
            [
                [1, 0, -1] * evses_per_phase,
                [-1, 1, 0] * evses_per_phase,
                [0, -1, 1] * evses_per_phase,
            ],
            dtype=float,
        ),
        "constraint_limits": np.repeat(limit, 3).astype(float),
        "phases": np.array([0, -120, 120] * evses_per_phase, dtype=float),
        "voltages": np.repeat(208, n),
        "constraint_ids": ["A", "B", "C"],
        "station_ids": [f"{i}" for i in range(n)],
        "max_pilot": np.repeat(max_pilot, n),
        "min_pilot": np.repeat(min_pilot, n),
        "allowable_pilots": allowable_pilots,
        "is_continuous": is_continuous,
    }


# ------------------------------------------------------------------------------ sites
def caltech_acn_infrastructure(voltage=208, transformer_cap=150) -> Dict:
    """54 EVSEs on a delta-connected 3-phase transformer: 26 on AB (30 deg; incl. an
    8-EVSE AeroVironment pod and an 8-EVSE ClipperCreek pod), 14 on BC (-90 deg), 14
    on CA (150 deg).  Constraints: 3 secondary line currents, 3 primary line
    currents (1/4 turns ratio, differences of secondary phase currents), 2 pods."""
    cc_pod = ["CA-322", "CA-493", "CA-496", "CA-320", "CA-495", "CA-321", "CA-323", "CA-494"]
    av_pod = ["CA-324", "CA-325", "CA-326", "CA-327", "CA-489", "CA-490", "CA-491", "CA-492"]
    ab = [f"CA-{i}" for i in [308, 508, 303, 513, 310, 506, 316, 500, 318, 498]] + av_pod + cc_pod
    bc = [f"CA-{i}" for i in [304, 512, 305, 511, 313, 503, 311, 505, 317, 499, 148, 149, 212, 213]]
    ca = [f"CA-{i}" for i in [307, 509, 309, 507, 306, 510, 315, 501, 319, 497, 312, 504, 314, 502]]
    ids = ab + bc + ca
    n = len(ids)
    idx = {s: i for i, s in enumerate(ids)}
    phases = np.array([30.0] * len(ab) + [-90.0] * len(bc) + [150.0] * len(ca))

    def cur(names):
        v = np.zeros(n)
        for s in names:
            v[idx[s]] = 1.0
        return v

    i3a, i3b, i3c = cur(ab), cur(bc), cur(ca)
    prim = transformer_cap * 1000 / 3 / 277
    sec = transformer_cap * 1000 / 3 / 120
    rows = [i3a, i3b, i3c, 0.25 * (i3a - i3c), 0.25 * (i3b - i3a), 0.25 * (i3c - i3b), cur(av_pod), cur(cc_pod)]
    limits = [sec, sec, sec, prim, prim, prim, 80.0, 80.0]
    names = ["Secondary A", "Secondary B", "Secondary C", "Primary A", "Primary B", "Primary C", "AV Pod", "CC Pod"]
    pilots = []
    for s in ids:
        if s in av_pod or s in ab[:10] or s in bc or s in ca:
            # AeroVironment: 0, 6..32 in 1 A steps
            pilots.append(np.array([0] + list(range(6, 33)), dtype=float))
        else:
            # ClipperCreek: discrete levels
            pilots.append(np.array([0, 8, 16, 24, 32], dtype=float))
    return {
        "constraint_matrix": np.array(rows),
        "constraint_limits": np.array(limits),
        "phases": phases,
        "voltages": np.repeat(float(voltage), n),
        "constraint_ids": names,
        "station_ids": ids,
        "max_pilot": np.repeat(32.0, n),
        "min_pilot": np.array([p[1] for p in pilots]),
        "allowable_pilots": pilots,
        "is_continuous": np.array([len(p) > 6 for p in pilots]),
    }


def hierarchical_three_phase_network(num_evses=1000, evses_per_pod=20, pods_per_panel=5, voltage=208.0, seed=0) -> Dict:
    """Synthetic large site (config C5): EVSEs grouped into single-phase-pair pods
    (line-to-line on AB/BC/CA in rotation), pods into panels, panels into one
    transformer.  Rows: one per pod (same-phase sum, limit 0.6 * 32 A * pod size),
    three line currents per panel, three secondary and three primary line currents
    for the transformer.  M = pods + 3*panels + 6."""
    n = num_evses
    n_pods = (n + evses_per_pod - 1) // evses_per_pod
    n_panels = (n_pods + pods_per_panel - 1) // pods_per_panel
    pod_of = np.arange(n) // evses_per_pod
    panel_of = pod_of // pods_per_panel
    phase_id = pod_of % 3
    phases = np.array([30.0, -90.0, 150.0])[phase_id]
    rows, limits, names = [], [], []
    for p in range(n_pods):
        v = (pod_of == p).astype(float)
        rows.append(v)
        limits.append(0.6 * 32.0 * v.sum())
        names.append(f"pod{p}")
    def line_rows(mask, tag, frac, scale=1.0):
        a = (mask & (phase_id == 0)).astype(float)
        b = (mask & (phase_id == 1)).astype(float)
        c = (mask & (phase_id == 2)).astype(float)
        tot = mask.sum()
        for nm, v in (("a", a - c), ("b", b - a), ("c", c - b)):
            rows.append(scale * v)
            limits.append(scale * frac * 32.0 * tot / 3 * np.sqrt(3))
            names.append(f"{tag}.{nm}")
    for q in range(n_panels):
        line_rows(panel_of == q, f"panel{q}", 0.5)
    allm = np.ones(n, dtype=bool)
    line_rows(allm, "sec", 0.4)
    line_rows(allm, "pri", 0.38, scale=0.25)
    pilots = [np.array([0] + list(range(6, 33)), dtype=float) for _ in range(n)]
    return {
        "constraint_matrix": np.array(rows),
        "constraint_limits": np.array(limits),
        "phases": phases,
        "voltages": np.repeat(float(voltage), n),
        "constraint_ids": names,
        "station_ids": [f"EV-{i:04d}" for i in range(n)],
        "max_pilot": np.repeat(32.0, n),
        "min_pilot": np.repeat(6.0, n),
        "allowable_pilots": pilots,
        "is_continuous": np.ones(n, dtype=bool),
    }


# -------------------------------------------------------------------------- workloads
def sce_tou_prices(T, period=5, start_hour=6.0, noise=0.0, rng=None):
    """Three-level SCE-like TOU tariff ($/kWh): off-peak 0.056 (23-8h), mid 0.092,
    on-peak 0.267 (12-18h); the horizon starts at ``start_hour``."""
    h = (start_hour + np.arange(T) * period / 60.0) % 24
    p = np.where((h >= 12) & (h < 18), 0.267, np.where((h >= 8) & (h < 23), 0.092, 0.056))
    if noise > 0:
        p = p * (1 + noise * (2 * rng.random(T) - 1))
    return p


def _deliverable_kwh(dur, max_rate, voltage, period):
    return dur * max_rate * voltage / 1000.0 * period / 60.0


def config_c1(seed=0, n=30, T=144, period=5) -> Dict:
    """C1: single-phase N=30, M=1, limit 32*N/3, 30 sessions, arrival 0,
    departure U{36..T} (max forced to T), demand U[2,20] kWh clipped to 90% of
    deliverable."""
    rng = np.random.default_rng(seed)
    dep = rng.integers(36, T + 1, size=n)
    dep[rng.integers(0, n)] = T
    dem = rng.uniform(2, 20, size=n)
    dem = np.minimum(dem, 0.9 * _deliverable_kwh(dep, 32, 208, period))
    sessions = session_generator(n, [0] * n, dep.tolist(), dem.tolist(), dem.tolist(), [32] * n)
    infra = single_phase_single_constraint(n, 32 * n / 3)
    return {
        "active_sessions": sessions,
        "infrastructure_info": infra,
        "current_time": 0,
        "period": period,
    }


def config_c2(seed=0, T=288, period=5, infra: Optional[Dict] = None, price_noise=0.0) -> Dict:
    """C2/C3: CaltechACN-shaped site, 20..54 sessions (one per EVSE), arrivals
    U{0..96}, durations U{24..192} truncated at T (one session forced to end at T),
    demand U[2,20] kWh clipped to 90% of deliverable, TOU prices, demand charge
    15.51 $/kW, prev_peak U[0,100] A."""
    rng = np.random.default_rng(seed)
    if infra is None:
        infra = caltech_acn_infrastructure()
    n = len(infra["station_ids"])
    s = int(rng.integers(20, n + 1))
    stations = rng.permutation(n)[:s]
    arr = rng.integers(0, 97, size=s)
    dur = rng.integers(24, 193, size=s)
    dep = np.minimum(arr + dur, T)
    dep[int(rng.integers(0, s))] = T
    dem = rng.uniform(2, 20, size=s)
    dem = np.minimum(dem, 0.9 * _deliverable_kwh(dep - arr, 32, 208, period))
    sessions = session_generator(
        s,
        arr.tolist(),
        dep.tolist(),
        dem.tolist(),
        dem.tolist(),
        [32] * s,
        station_ids=[infra["station_ids"][i] for i in stations],
    )
    return {
        "active_sessions": sessions,
        "infrastructure_info": infra,
        "current_time": 0,
        "period": period,
        "prices": sce_tou_prices(T, period, noise=price_noise, rng=rng),
        "demand_charge": 15.51,
        "prev_peak": float(rng.uniform(0, 100)),
    }


def config_c5(seed=0, T=288, period=5, infra: Optional[Dict] = None, n=1000) -> Dict:
    """C5: 1000-EVSE hierarchical three-phase site, 60..100% occupancy."""
    rng = np.random.default_rng(seed)
    if infra is None:
        infra = hierarchical_three_phase_network(n)
    n = len(infra["station_ids"])
    s = int(rng.integers(int(0.6 * n), n + 1))
    stations = rng.permutation(n)[:s]
    arr = rng.integers(0, 97, size=s)
    dur = rng.integers(24, 193, size=s)
    dep = np.minimum(arr + dur, T)
    dep[int(rng.integers(0, s))] = T
    dem = rng.uniform(2, 20, size=s)
    dem = np.minimum(dem, 0.9 * _deliverable_kwh(dep - arr, 32, 208, period))
    sessions = session_generator(
        s, arr.tolist(), dep.tolist(), dem.tolist(), dem.tolist(), [32] * s,
        station_ids=[infra["station_ids"][i] for i in stations],
    )
    t = np.arange(T)
    ext = 150.0 + 100.0 * np.sin(2 * np.pi * (t / T - 0.25)) + 10.0 * rng.standard_normal(T)
    return {
        "active_sessions": sessions,
        "infrastructure_info": infra,
        "current_time": 0,
        "period": period,
        "prices": sce_tou_prices(T, period),
        "demand_charge": 15.51,
        "prev_peak": 0.0,
        "external_signal": ext,
    }


__all__ = [
    "session_generator",
    "single_phase_single_constraint",
    "three_phase_balanced_network",
    "caltech_acn_infrastructure",
    "hierarchical_three_phase_network",
    "sce_tou_prices",
    "config_c1",
    "config_c2",
    "config_c5",
]
