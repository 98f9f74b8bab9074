"""Postprocessing — same functions as reference adacharge/postprocessing.py ("pp.py")
and adacharge/utils.py, executed by float64 CUDA kernels (csrc/acb_post.cu).

Scalar helpers (floor_to_set, ceil_to_set, increment_in_set; pp.py:10-74) are host
Python: they are one-element utilities, the device twins live inside the kernels.
"""
from __future__ import annotations

from typing import List

import numpy as np

from .interface import Interface, SessionInfo, InfrastructureInfo
from . import engine


def _member(allowable_set, index: int):
    """allowable_set[index] with the index clipped to the ends of the set."""
    return allowable_set[min(max(index, 0), len(allowable_set) - 1)]


def floor_to_set(x: float, allowable_set: np.ndarray, eps=0.05):
    """Round x down into allowable_set (a member within eps above x counts as reached); an exact member is returned
    unchanged; results are clipped to the ends of the set.  Same contract as reference postprocessing.py:10-31
    (device twin: floor_to_set in csrc/acb_post.cu)."""
    below = int(np.searchsorted(np.asarray(allowable_set), x + eps, side="left"))  # members [0, below) are < x + eps
    if below < len(allowable_set) and allowable_set[below] == x:
        return x
    return _member(allowable_set, below - 1)


def ceil_to_set(x: float, allowable_set: np.ndarray, eps=0.05):
    """Round x up into allowable_set (a member within eps below x counts as reached); reference postprocessing.py:34-55."""
    upto = int(np.searchsorted(np.asarray(allowable_set), x - eps, side="right"))  # members [0, upto) are <= x - eps
    if upto > 0 and allowable_set[upto - 1] == x:
        return x
    return _member(allowable_set, upto)


def increment_in_set(x: float, allowable_set: np.ndarray):
    """Smallest member strictly above x, or the largest member if there is none; reference postprocessing.py:58-74."""
    return _member(allowable_set, int(np.searchsorted(np.asarray(allowable_set), x, side="right")))


def _site(infrastructure, network=True) -> engine.Site:
    """Device constants for postprocessing.  The projections only read max_pilot / allowable_pilots (network=False:
    works for infrastructures without phases, like the reference); reallocation and the feasibility check need the
    SOC rows.  Any site already built for this infrastructure (e.g. by the solve) is reused."""
    return engine.get_post_site(infrastructure, network)


def _back(out, like):
    a = out.cpu().numpy()
    dt = np.asarray(like).dtype
    return a.astype(dt) if dt != np.float64 else a


def project_into_continuous_feasible_pilots(rates: np.ndarray, infrastructure: InfrastructureInfo):
    """Clip every rate into [0, max_pilot of its EVSE]; dtype preserved (pp.py:77-94)."""
    return _back(engine.project_continuous(_site(infrastructure, network=False), rates), rates)


def project_into_discrete_feasible_pilots(rates: np.ndarray, infrastructure: InfrastructureInfo):
    """floor_to_set(., allowable_pilots[i], eps=0.05) then max(., 0); dtype preserved (pp.py:97-118)."""
    return _back(engine.project_discrete(_site(infrastructure, network=False), rates), rates)


def _session_arrays(active_sessions, infrastructure, interface):
    S = max(1, len(active_sessions))
    row = np.zeros((1, S), np.int32)
    st = np.zeros((1, S), np.int32)
    ramp = np.zeros((1, S))
    mx0 = np.zeros((1, S))
    for k, s in enumerate(active_sessions):
        row[0, k] = infrastructure.get_station_index(s.station_id)
        st[0, k] = s.arrival_offset
        if s.arrival_offset == 0:  # only these are read (pp.py:157, 229)
            ramp[0, k] = interface.remaining_amp_periods(s)
            mx0[0, k] = np.asarray(s.max_rates, dtype=float).ravel()[0]
    return np.array([len(active_sessions)], np.int32), row, st, ramp, mx0


def index_based_reallocation(rates: np.ndarray, active_sessions: List[SessionInfo], infrastructure: InfrastructureInfo,
                             peak_limit: float, sort_fn, interface: Interface):
    """Greedy round-robin increment of the first period up to peak_limit, order from
    sort_fn; ``rates`` is modified in place and returned (pp.py:121-186)."""
    ns, row, st, ramp, mx0 = _session_arrays(active_sessions, infrastructure, interface)
    sorted_sessions = sort_fn(active_sessions, interface)
    pos = {id(s): k for k, s in enumerate(active_sessions)}
    order = np.zeros_like(row)
    order[0, : len(sorted_sessions)] = [pos[id(s)] for s in sorted_sessions]
    out = engine.reallocate(_site(infrastructure), 0, np.asarray(rates)[None], ns, row, st, ramp, mx0, order, np.array([float(peak_limit)]))
    rates[...] = _back(out[0], rates)
    return rates


def diff_based_reallocation(rates: np.ndarray, active_sessions: List[SessionInfo], infrastructure: InfrastructureInfo,
                            interface: Interface):
    """Round down, then re-add first-period capacity ordered by the largest rounding
    loss; returns a new array (pp.py:189-258)."""
    ns, row, st, ramp, mx0 = _session_arrays(active_sessions, infrastructure, interface)
    out = engine.reallocate(_site(infrastructure), 1, np.asarray(rates)[None], ns, row, st, ramp, mx0)
    return _back(out[0], rates)


def infrastructure_constraints_feasible(rates, infrastructure: InfrastructureInfo):
    """All SOC line currents <= limit + 1e-7, for a vector (N) or every column of a
    matrix (utils.py:5-12)."""
    r = np.asarray(rates, dtype=np.float64)
    site = _site(infrastructure)
    if r.ndim == 1:
        return bool(engine.constraints_feasible(site, r[None, :, None], 0)[0].item())
    return all(bool(engine.constraints_feasible(site, r[None], c)[0].item()) for c in range(r.shape[1]))


__all__ = [
    "floor_to_set", "ceil_to_set", "increment_in_set", "project_into_continuous_feasible_pilots",
    "project_into_discrete_feasible_pilots", "index_based_reallocation", "diff_based_reallocation",
    "infrastructure_constraints_feasible",
]
