"""ORACLE (test infrastructure, never the product path): numpy restatement of the
reference postprocessing, reference adacharge/postprocessing.py ("pp.py") and
adacharge/utils.py.  Pinned: tests/test_oracle_postprocessing.py checks it against
every known answer in the reference's tests/test_postprocessing.py and against golden
vectors produced by the reference's own pp.py (tests/golden/make_golden.py).

Objects passed in need only the attributes listed in SURVEY.md §8(b).
"""
from __future__ import annotations

import numpy as np


def floor_to_set(x, allowable, eps=0.05):
    """pp.py:10-31.  Largest element strictly below x+eps, except that an exact member
    is returned as is; clipped to the ends of the set."""
    a = np.asarray(allowable)
    pos = int(np.searchsorted(a, x + eps, side="left"))
    if pos < len(a) and x == a[pos]:
        return x
    return a[min(max(pos - 1, 0), len(a) - 1)] if pos > 0 else a[0]


def ceil_to_set(x, allowable, eps=0.05):
    """pp.py:34-55."""
    a = np.asarray(allowable)
    pos = int(np.searchsorted(a, x - eps, side="right"))
    if pos > 0 and x == a[pos - 1]:
        return x
    return a[min(pos, len(a) - 1)]


def increment_in_set(x, allowable):
    """pp.py:58-74.  Next strictly larger element, clipped to the ends."""
    a = np.asarray(allowable)
    pos = int(np.searchsorted(a, x, side="right"))
    return a[min(pos, len(a) - 1)]


def project_into_continuous_feasible_pilots(rates, infrastructure):
    """pp.py:77-94: min with max_pilot per row, then max with 0; dtype follows numpy's rules."""
    r = np.array(rates, copy=True)
    mp = np.asarray(infrastructure.max_pilot)
    for i in range(infrastructure.num_stations):
        r[i] = np.minimum(rates[i], mp[i])
    return np.maximum(r, 0)


def project_into_discrete_feasible_pilots(rates, infrastructure):
    """pp.py:97-118: element-wise floor_to_set(eps=0.05) into the EVSE's allowable pilots,
    then max with 0."""
    r = np.array(rates, copy=True)
    n, T = r.shape
    for i in range(infrastructure.num_stations):
        a = np.array(infrastructure.allowable_pilots[i])
        for t in range(T):
            r[i, t] = floor_to_set(rates[i, t], a, eps=0.05)
    return np.maximum(r, 0)


def infrastructure_constraints_feasible(rates, infrastructure):
    """utils.py:5-12: every SOC line current <= limit + 1e-7 (vector or matrix input)."""
    ph = np.deg2rad(infrastructure.phases)
    for j, v in enumerate(infrastructure.constraint_matrix):
        a = np.stack([v * np.cos(ph), v * np.sin(ph)])
        cur = np.linalg.norm(a @ rates, axis=0)
        if not np.all(cur <= infrastructure.constraint_limits[j] + 1e-7):
            return False
    return True


def _first_period_caps(active_sessions, infrastructure, interface):
    """pp.py:152-164 / 224-236: EVSEs whose session starts now, and their caps."""
    n = infrastructure.num_stations
    active = np.zeros(n, dtype=bool)
    ub = np.zeros(n)
    for s in active_sessions:
        if s.arrival_offset == 0:
            i = infrastructure.station_ids.index(s.station_id)
            active[i] = True
            ub[i] = min(interface.remaining_amp_periods(s), s.max_rates[0], infrastructure.max_pilot[i])
    return active, ub


def _greedy_first_period(col_owner, order, active, ub, peak_limit, infrastructure, guard=True):
    """pp.py:166-185 / 238-257: round-robin over `order`, raising one EVSE one pilot step
    at a time while the column stays under peak_limit, the EVSE under its cap and the
    network feasible.  `guard` stops an EVSE whose increment makes no progress (the
    reference would loop forever there)."""
    if len(order) == 0:
        return
    k = 0
    idle = 0
    while active.any():
        i = order[k]
        k = (k + 1) % len(order)
        if not active[i]:
            idle += 1
            if idle >= len(order):
                break
            continue
        idle = 0
        if col_owner[i, 0] >= ub[i]:
            active[i] = False
            continue
        trial = np.array(col_owner[:, 0], copy=True)
        trial[i] = increment_in_set(col_owner[i, 0], infrastructure.allowable_pilots[i])
        ok = (np.sum(trial) <= peak_limit and trial[i] <= ub[i]
              and infrastructure_constraints_feasible(trial, infrastructure))
        if guard and trial[i] == col_owner[i, 0]:
            ok = False
        if ok:
            col_owner[:, 0] = trial
        else:
            active[i] = False


def index_based_reallocation(rates, active_sessions, infrastructure, peak_limit, sort_fn, interface):
    """pp.py:121-186; mutates and returns `rates`."""
    order = [infrastructure.get_station_index(s.station_id) for s in sort_fn(active_sessions, interface)]
    active, ub = _first_period_caps(active_sessions, infrastructure, interface)
    _greedy_first_period(rates, order, active, ub, peak_limit, infrastructure)
    return rates


def diff_based_reallocation(rates, active_sessions, infrastructure, interface):
    """pp.py:189-258; returns a new rounded + reallocated array."""
    init = rates[:, 0]
    peak_limit = init.sum()
    rounded = project_into_discrete_feasible_pilots(rates, infrastructure)

    def loss(s):
        i = infrastructure.get_station_index(s.station_id)
        return -(init[i] - rounded[i, 0])

    order = [infrastructure.get_station_index(s.station_id) for s in sorted(active_sessions, key=loss)]
    active, ub = _first_period_caps(active_sessions, infrastructure, interface)
    _greedy_first_period(rounded, order, active, ub, peak_limit, infrastructure)
    return rounded
