"""Stand-ins for the acnportal.algorithms pieces the reference adapter imports
(reference adacharge/adacharge.py:1-10): BaseAlgorithm and the three preprocessing
helpers.  [acnportal, recalled]; real acnportal classes are used when importable.
These run on the host before every solve."""
from __future__ import annotations

from copy import deepcopy

import numpy as np

try:  # pragma: no cover
    from acnportal.algorithms import (  # type: ignore
        BaseAlgorithm, apply_upper_bound_estimate, apply_minimum_charging_rate, enforce_pilot_limit,
    )
except Exception:  # noqa: BLE001

    class BaseAlgorithm:
        def __init__(self):
            self._interface = None
            self.max_recompute = None

        @property
        def interface(self):
            if self._interface is not None:
                return self._interface
            raise ValueError("No interface has been registered yet. Please call register_interface prior to using the algorithm.")

        def register_interface(self, interface):
            self._interface = interface

        def schedule(self, active_sessions):
            raise NotImplementedError

        def run(self):
            return self.schedule(self.interface.active_sessions())

    def _expand(sessions):
        out = deepcopy(sessions)
        for s in out:
            rt = s.remaining_time
            if np.isscalar(s.max_rates):
                s.max_rates = np.full(rt, float(s.max_rates))
            if np.isscalar(s.min_rates):
                s.min_rates = np.full(rt, float(s.min_rates))
            s.max_rates = np.array(s.max_rates, dtype=float)
            s.min_rates = np.array(s.min_rates, dtype=float)
        return out

    def _reconcile(session, choose_min=True):
        mask = session.max_rates < session.min_rates
        if choose_min:
            session.max_rates[mask] = session.min_rates[mask]
        else:
            session.min_rates[mask] = session.max_rates[mask]
        return session

    def enforce_pilot_limit(active_sessions, infrastructure):
        new = _expand(active_sessions)
        for s in new:
            i = infrastructure.get_station_index(s.station_id)
            s.max_rates = np.minimum(s.max_rates, infrastructure.max_pilot[i])
        return new

    def apply_upper_bound_estimate(ub_estimator, active_sessions):
        new = _expand(active_sessions)
        upper = ub_estimator.get_maximum_rates(active_sessions)
        for j, s in enumerate(new):
            s.max_rates = np.minimum(s.max_rates, upper.get(s.session_id, float("inf")))
            new[j] = _reconcile(s)
        return new

    def apply_minimum_charging_rate(active_sessions, infrastructure, override=float("inf")):
        """Sessions in arrival order keep their EVSE's minimum pilot in the first period while the network stays
        feasible (acnportal preprocessing, called at ada.py:149-150).  The sequential feasibility-guarded greedy
        runs on the device in one launch (acb_min_rate_admission)."""
        from . import engine

        queue = _expand(sorted(active_sessions, key=lambda x: x.arrival))
        if not queue:
            return queue
        rows = np.array([[infrastructure.get_station_index(s.station_id) for s in queue]], dtype=np.int32)
        tries = np.array([[min(infrastructure.min_pilot[i], override) for i in rows[0]]], dtype=np.float64)
        site = engine.get_site(infrastructure, "SOC", False, False)
        admitted = engine.min_rate_admission(site, [len(queue)], rows, tries)[0]
        for j, s in enumerate(queue):
            if admitted[j]:
                s.min_rates[0] = max(tries[0, j], s.min_rates[0])
                queue[j] = _reconcile(s)
            else:
                s.min_rates[0] = 0
                s.max_rates[0] = 0
        return queue
