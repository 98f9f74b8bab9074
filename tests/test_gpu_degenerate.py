"""Degenerate closed-loop instances captured from the 1024-site replay (tests/golden/make_replay_degenerate.py):
one EV whose remaining energy is exactly what the site's previous peak lets it draw, so the sunk
demand charge w*p0 cancels the energy revenue (|P| << |terms|) and the LP optimum is a vertex a
few mA away from the flat schedule.  Each instance is a 1-EV LP, so the exact optimum comes from
scipy's HiGHS here; the solver's certificate must bound the true suboptimality."""
import os

import numpy as np
import pytest

from adacharge_b200 import _cabi, engine
from adacharge_b200.generators import caltech_acn_infrastructure
from adacharge_b200.interface import InfrastructureInfo

pytestmark = pytest.mark.gpu
FIX = os.path.join(os.path.dirname(__file__), "golden", "replay_degenerate_instances.npz")
TP, N = 160, 54


def _site():
    infra = caltech_acn_infrastructure()
    info = InfrastructureInfo(np.asarray(infra["constraint_matrix"]), np.asarray(infra["constraint_limits"]), np.asarray(infra["phases"]),
                              np.asarray(infra["voltages"], float), infra["constraint_ids"], infra["station_ids"],
                              np.asarray(infra["max_pilot"]), np.asarray(infra["min_pilot"]))
    return engine.get_site(info, "SOC", False, True), np.asarray(infra["voltages"], float) / 1e3


def _lp_optimum(h, j, k):
    """min c.r + w*tau  s.t. k r_t <= tau, tau >= p0, sum r <= E, 0 <= r <= max  (one EV, network limits slack)."""
    from scipy.optimize import linprog

    T, i = int(h["T"][j]), int(h["sess_row"][j, 0])
    c = h["alpha"][j, :T].astype(float) + k[i] * h["beta"][j, :T].astype(float)
    w, p0, E = float(h["peak_w"][j]), float(h["peak_p0"][j]), float(h["sess_energy"][j, 0])
    ub = float(h["max_rates"][-(int(h["sess_rate_off"][j, 0]) + 1)])
    A = np.zeros((T + 1, T + 1))
    A[np.arange(T), np.arange(T)] = k[i]
    A[:T, T] = -1
    A[T, :T] = 1
    res = linprog(np.r_[c, w], A_ub=A, b_ub=np.r_[np.zeros(T), E], bounds=[(0, ub)] * T + [(p0, None)], method="highs")
    assert res.status == 0
    return res.fun, c, w, p0, i, T


def _objective(h, j, k, rates):
    _, c, w, p0, i, T = _lp_optimum(h, j, k)
    r = rates[j, i, :T].astype(float)
    lin, pk = float(c @ r), w * max(float((k[i] * r).max()), p0)
    return lin + pk, abs(lin) + abs(pk)


def _solve(h, **kw):
    site, k = _site()
    pb = engine.PackedBatch.from_arrays(site, h, TP, N).upload().solve(_cabi.default_options(**kw))
    return pb.rates.cpu().numpy(), pb.status.cpu().numpy(), pb.iters.cpu().numpy(), pb.stats.cpu().numpy(), k


def test_term_magnitude_tolerance_certifies_degenerate_instances(require_gpu):
    h = dict(np.load(FIX))
    rates, st, it, sx, k = _solve(h, term_floor=1.0)
    assert (st == _cabi.ACB_SOLVED).all(), (st, it, sx[:, 2])
    assert it.max() <= 5000
    for j in range(len(st)):
        opt, *_ = _lp_optimum(h, j, k)
        P, mag = _objective(h, j, k, rates)
        assert P >= opt - 1e-6 * mag  # feasible schedule cannot beat the optimum
        assert P - opt <= 1e-4 * mag, (j, P, opt, mag)
        # the certified gap (relative to the same scale) bounds the true suboptimality
        assert P - opt <= sx[j, 2] * mag * (1 + 1e-3) + 1e-6 * mag, (j, P - opt, sx[j, 2] * mag)


def test_stall_exit_bounds_the_work_and_keeps_the_schedule(require_gpu):
    """Relative to |P| alone (term_floor 0) these instances cannot be certified in fp32; the stall exit ends
    them early with ACB_MAX_ITER, a feasible schedule and an honest gap."""
    h = dict(np.load(FIX))
    rates, st, it, sx, k = _solve(h, term_floor=0.0, stall_exit=40)
    assert set(st.tolist()) <= {_cabi.ACB_SOLVED, _cabi.ACB_MAX_ITER}
    assert (st == _cabi.ACB_MAX_ITER).any()
    assert it.max() <= 8000  # not the 20 000 cap
    for j in range(len(st)):
        opt, c, w, p0, i, T = _lp_optimum(h, j, k)
        P, mag = _objective(h, j, k, rates)
        assert -1e-6 * mag <= P - opt <= 1e-3 * mag, (j, P, opt)
        assert rates[j, i, :T].sum() <= float(h["sess_energy"][j, 0]) * (1 + 1e-5)
        assert rates[j].min() >= 0 and rates[j].max() <= 32.0 + 1e-4


@pytest.mark.parametrize("refine", [0, 2])
def test_dual_bound_is_valid_with_and_without_refinement(require_gpu, refine):
    h = dict(np.load(FIX))
    rates, st, it, sx, k = _solve(h, term_floor=1.0, dual_refine=refine, max_iter=600)
    for j in range(len(st)):
        opt, *_ = _lp_optimum(h, j, k)
        P, mag = _objective(h, j, k, rates)
        # whatever the status, the reported gap (P - D) / scale must cover the true suboptimality: D <= optimum
        assert P - opt <= sx[j, 2] * mag * (1 + 1e-3) + 2e-6 * mag, (j, refine, st[j], P - opt, sx[j, 2] * mag)
