"""A tiny tracer for user-defined objective components.

In the reference an ``ObjectiveComponent.function`` is any callable that builds a cvxpy expression from the ``rates``
variable (reference adacharge/adaptive_charging_optimization.py:200-218); cvxpy then canonicalises it.  Without
cvxpy the device solver needs the component in its packed form (``kernel_spec``).  For callables that are *numeric*
-- ``f(rates: ndarray, infrastructure, interface, **kwargs) -> float``, like the built-in twins in
``adaptive_charging_optimization`` -- this module recovers that form by probing:

    f(r) = const - sum_it (alpha_t + k_i beta_t) r_it - qd sum r_it^2 - gamma sum_t (sum_i k_i r_it)^2

(k_i = V_i / 1000; a load-flattening external signal only shifts beta_t and the constant).  The probes are finite
differences, exact for quadratics; the fitted model is then checked against ``f`` at random schedules, and a
component that does not fit (a norm, a max, cross-period coupling, EVSE-specific weights) is rejected loudly.
"""
from __future__ import annotations

from typing import Dict

import numpy as np

_CACHE: Dict = {}


class NotTraceable(TypeError):
    pass


def trace_component(fn, infrastructure, interface, T: int, **kwargs) -> dict:
    """Kernel spec (the dict the built-in ``_spec_*`` functions return) of ONE unit of the maximised component ``fn``."""
    N = len(infrastructure.station_ids)
    k = np.asarray(infrastructure.voltages, dtype=float) / 1e3

    def f(R):
        try:
            return float(fn(R, infrastructure, interface, **kwargs))
        except Exception as e:  # a cvxpy-style callable handed a numpy array, a missing kwarg, ...
            raise NotTraceable(f"objective component {getattr(fn, '__name__', fn)!r} cannot be evaluated on a numpy rates matrix "
                               f"({type(e).__name__}: {e}); the device solver needs a numeric callable or a `kernel_spec`") from e

    Z = np.zeros((N, T))
    f0 = f(Z)
    # two probe EVSEs: different kW/A if the site has them (separates alpha from beta), else any two
    i1 = 0
    diff = np.nonzero(np.abs(k - k[0]) > 1e-12 * abs(k[0]))[0]
    i2 = int(diff[0]) if len(diff) else (1 if N > 1 else 0)
    d = 1.0

    def first_second(i, t):
        R = Z.copy()
        R[i, t] = d
        a = f(R)
        R[i, t] = 2 * d
        b = f(R)
        return (4 * a - b - 3 * f0) / (2 * d), (b - 2 * a + f0) / (d * d)  # gradient and second derivative at 0

    g1 = np.zeros(T); g2 = np.zeros(T); h1 = np.zeros(T)
    for t in range(T):
        g1[t], h1[t] = first_second(i1, t)
        g2[t], _ = first_second(i2, t)
    # cross second difference inside one period -> aggregate quadratic; across periods -> must vanish
    gamma = 0.0
    if N > 1:
        j = i2 if i2 != i1 else 1
        R = Z.copy(); R[i1, 0] = d; R[j, 0] = d
        fa = f(R)
        R1 = Z.copy(); R1[i1, 0] = d
        R2 = Z.copy(); R2[j, 0] = d
        cross = (fa - f(R1) - f(R2) + f0) / (d * d)      # = -2 gamma k_i k_j
        gamma = -cross / (2 * k[i1] * k[j])
    qd = -(h1[0] + 2 * gamma * k[i1] ** 2) / 2           # h_ii = -2 qd - 2 gamma k_i^2
    # linear part: -g_it = alpha_t + k_i beta_t
    if abs(k[i2] - k[i1]) > 1e-12 * abs(k[i1]):
        beta = -(g2 - g1) / (k[i2] - k[i1])
        alpha = -g1 - k[i1] * beta
    else:
        alpha, beta = -g1, np.zeros(T)
    spec = dict(alpha=alpha, beta=beta, qd=float(qd), gamma=float(gamma), ext=np.zeros(T))
    # validation against the callable itself
    rng = np.random.default_rng(0)
    scale = max(1.0, abs(f0))
    for _ in range(4):
        R = rng.uniform(0, 8, size=(N, T))
        u = k @ R
        model = f0 - ((alpha[None, :] + k[:, None] * beta[None, :]) * R).sum() - qd * (R * R).sum() - gamma * (u * u).sum()
        val = f(R)
        scale = max(scale, abs(val))
        if abs(val - model) > 1e-7 * scale:
            raise NotTraceable(
                f"objective component {getattr(fn, '__name__', fn)!r} is not of the form the device solver packs "
                "(linear with period and kW-per-A weights, plus a uniform diagonal quadratic, plus a quadratic in the aggregate power): "
                f"model {model:.9g} vs function {val:.9g} at a random schedule; give it a `kernel_spec` or use the built-ins")
    if qd < -1e-12 * scale or gamma < -1e-12 * scale:
        raise NotTraceable(f"objective component {getattr(fn, '__name__', fn)!r} is convex, not concave, in the rates (qd {qd:.3g}, gamma {gamma:.3g})")
    spec["qd"], spec["gamma"] = max(spec["qd"], 0.0), max(spec["gamma"], 0.0)
    if spec["gamma"] == 0.0:
        del spec["gamma"], spec["ext"]
    if spec["qd"] == 0.0:
        del spec["qd"]
    return spec


def traced_spec(fn):
    """``kernel_spec`` for ``fn`` obtained by tracing (cached per horizon and kwargs)."""

    def spec(infrastructure, interface, T, **kwargs):
        key = (id(fn), id(infrastructure), T, getattr(interface, "current_time", None), tuple(sorted((k_, id(v)) for k_, v in kwargs.items())))
        if key not in _CACHE:
            if len(_CACHE) > 64:
                _CACHE.clear()
            _CACHE[key] = trace_component(fn, infrastructure, interface, T, **kwargs)
        return _CACHE[key]

    return spec


__all__ = ["trace_component", "traced_spec", "NotTraceable"]
