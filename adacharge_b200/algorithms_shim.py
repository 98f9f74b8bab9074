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
        from .postprocessing import infrastructure_constraints_feasible

        queue = _expand(sorted(active_sessions, key=lambda x: x.arrival))
        rates = np.zeros(len(infrastructure.station_ids))
        for j, s in enumerate(queue):
            i = infrastructure.get_station_index(s.station_id)
            rates[i] = min(infrastructure.min_pilot[i], override)
            if infrastructure_constraints_feasible(rates, infrastructure):
                s.min_rates[0] = max(rates[i], s.min_rates[0])
                queue[j] = _reconcile(s)
            else:
                rates[i] = 0
                s.min_rates[0] = 0
                s.max_rates[0] = 0
        return queue
