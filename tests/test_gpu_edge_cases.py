"""Edge cases of the solve path: ragged / degenerate inputs the reference accepts."""
import numpy as np
import pytest

import adacharge_b200 as ab
from adacharge_b200 import engine
from adacharge_b200.generators import session_generator, single_phase_single_constraint, three_phase_balanced_network, config_c2
from oracle import mpc

pytestmark = pytest.mark.gpu
QC = [ab.ObjectiveComponent(ab.quick_charge)]


def _iface(sessions, infra, **kw):
    return ab.TestingInterface({"active_sessions": sessions, "infrastructure_info": infra, "current_time": 0, "period": 5, **kw})


def _check(R, S, I, iface, spec, **kw):
    Ro = mpc.solve_mpc(spec, S, I, iface, **kw)
    f, fo = (mpc.evaluate_objective(X, spec, I, iface, S) for X in (R, Ro))
    assert abs(f - fo) <= 1e-4 * max(abs(fo), 1e-9), (f, fo)
    v = mpc.violations(R, S, I, iface, kw.get("constraint_type", "SOC"), kw.get("peak_limit"))
    # bounds are held in float32 on the device: a non-representable bound may be off by one ulp (< 1e-5 A)
    assert v["lb"] <= 1e-5 and v["ub"] <= 1e-5 and v["energy"] <= 1e-4 and v["infrastructure_rel"] <= 1e-5, v


def test_odd_horizon_and_single_session(require_gpu):
    s = session_generator(1, [3], [40], [6.0], [6.0], [32], station_ids=["4"])
    iface = _iface(s, single_phase_single_constraint(7, 20))
    S, I = iface.active_sessions(), iface.infrastructure_info()
    R = ab.AdaptiveChargingOptimization(QC, iface).solve(S, I)
    assert R.shape == (7, 40) and not R[[0, 1, 2, 3, 5, 6]].any() and not R[4, :3].any()
    _check(R, S, I, iface, [("quick_charge", 1, {})])


def test_zero_remaining_demand_and_zero_length_session(require_gpu):
    # one EV already full, one whose remaining_time is 0 (departure == current time), one normal
    s = session_generator(3, [0, 0, 0], [20, 0, 30], [5.0, 5.0, 5.0], [0.0, 2.0, 5.0], [32, 32, 32])
    iface = _iface(s, single_phase_single_constraint(3, 40))
    S, I = iface.active_sessions(), iface.infrastructure_info()
    R = ab.AdaptiveChargingOptimization(QC, iface).solve(S, I)
    assert not R[0].any() and not R[1].any() and R[2].sum() > 0
    assert abs(R[2].sum() * 208 * 5 / 6e4 - 5.0) < 1e-3


def test_ragged_rate_arrays_with_conflicting_bounds(require_gpu):
    # per-period min/max arrays; a min above the max must win (aco.py:75)
    s = session_generator(2, [0, 2], [10, 12], [3.0, 3.0], [3.0, 3.0], [32, 32])
    s[0]["max_rates"] = np.array([32, 32, 8, 8, 8, 32, 32, 32, 32, 32], dtype=float)
    s[0]["min_rates"] = np.array([0, 0, 10, 0, 0, 0, 0, 0, 0, 0], dtype=float)
    s[1]["max_rates"] = np.linspace(32, 6, 10)
    iface = _iface(s, single_phase_single_constraint(2, 50))
    S, I = iface.active_sessions(), iface.infrastructure_info()
    R = ab.AdaptiveChargingOptimization(QC, iface).solve(S, I)
    assert abs(R[0, 2] - 10.0) < 1e-4  # lb = ub = 10 there
    _check(R, S, I, iface, [("quick_charge", 1, {})])


def test_no_infrastructure_constraints(require_gpu):
    s = session_generator(3, [0, 1, 2], [12, 14, 9], [3.3] * 3, [3.3] * 3, [32] * 3)
    infra = single_phase_single_constraint(3, 64)
    infra["constraint_matrix"] = np.zeros((0, 0))
    infra["constraint_limits"] = np.zeros(0)
    infra["constraint_ids"] = []
    iface = _iface(s, infra)
    S, I = iface.active_sessions(), iface.infrastructure_info()
    R = ab.AdaptiveChargingOptimization([ab.ObjectiveComponent(ab.quick_charge), ab.ObjectiveComponent(ab.equal_share, 0.01)], iface).solve(S, I)
    need = 3.3 / (208 * 5 / 6e4)
    assert np.allclose(R.sum(axis=1), need, rtol=1e-4)


def test_linear_constraints_three_phase_tight(require_gpu):
    n = 12
    s = session_generator(n, [0] * n, [24] * n, [4.0] * n, [4.0] * n, [32] * n)
    iface = _iface(s, three_phase_balanced_network(n // 3, 40))
    S, I = iface.active_sessions(), iface.infrastructure_info()
    spec = [("quick_charge", 1, {}), ("equal_share", 1e-3, {})]
    for ct in ("LINEAR", "SOC"):
        aco = ab.AdaptiveChargingOptimization([ab.ObjectiveComponent(getattr(ab, a), b, c) for a, b, c in spec], iface, constraint_type=ct)
        R = aco.solve(S, I)
        _check(R, S, I, iface, spec, constraint_type=ct)


def test_idle_instances_inside_a_batch(require_gpu):
    """A batch may contain sites with a single short session next to full ones."""
    insts = []
    obj = [ab.ObjectiveComponent(ab.tou_energy_cost), ab.ObjectiveComponent(ab.total_energy, 0.3), ab.ObjectiveComponent(ab.demand_charge, 1 / 30)]
    for seed in range(3):
        d = config_c2(seed)
        if seed == 1:
            d["active_sessions"] = d["active_sessions"][:1]
        iface = ab.TestingInterface(d)
        S, I = iface.active_sessions(), iface.infrastructure_info()
        aco = ab.AdaptiveChargingOptimization(obj, iface)
        insts.append(aco.build_instance(S, I, None, iface.get_prev_peak()))
    pb = engine.PackedBatch(aco._site_for(I, insts[0]), insts).upload().solve()
    assert (pb.status.cpu().numpy() == 0).all()
    R = pb.rates.cpu().numpy()
    assert np.count_nonzero(np.abs(R[1]).sum(axis=1)) <= 1


def test_invalid_arguments_raise(require_gpu):
    iface = _iface(session_generator(1, [0], [5], [1.0], [1.0], [32]), single_phase_single_constraint(2, 30))
    S, I = iface.active_sessions(), iface.infrastructure_info()
    with pytest.raises(ValueError):
        ab.AdaptiveChargingOptimization(QC, iface, constraint_type="AFFINE").solve(S, I)
    I.phases = None
    with pytest.raises(ValueError):
        ab.AdaptiveChargingOptimization(QC, iface, constraint_type="SOC").solve(S, I)
    with pytest.raises(ValueError):
        ab.AdaptiveChargingOptimization(QC, iface).solve(iface.active_sessions(), iface.infrastructure_info(), peak_limit=[10.0, 10.0])


@pytest.mark.parametrize("Tp", [64, 128, 160, 288])
def test_every_padded_horizon_gives_the_same_schedule(require_gpu, Tp):
    """The kernel is instantiated per padded horizon (Tp = 32 Q); the result must not depend on the padding."""
    import adacharge_b200 as ab
    from adacharge_b200 import engine
    from adacharge_b200.generators import config_c1

    iface = ab.TestingInterface(config_c1(seed=5, n=9, T=40))
    S, I = iface.active_sessions(), iface.infrastructure_info()
    aco = ab.AdaptiveChargingOptimization([ab.ObjectiveComponent(ab.quick_charge), ab.ObjectiveComponent(ab.equal_share, 0.05)], iface)
    inst = aco.build_instance(S, I, None, 0)
    site = aco._site_for(I, inst)
    ref = aco.solve(S, I)  # smallest fitting horizon (64)
    pb = engine.PackedBatch(site, [inst], Tp=Tp).upload().solve(aco._options(inst))
    assert int(pb.status[0]) == 0
    got = pb.rates[0, :, : inst.T].cpu().numpy().astype(np.float64)
    assert (pb.rates[0, :, inst.T:] == 0).all()
    assert np.abs(got - ref).max() <= 2e-3  # same optimum (unique: strongly concave), different reduction trees
    # the general (streaming) path accepts the same horizons
    from adacharge_b200 import _cabi

    pg = engine.PackedBatch(site, [inst], Tp=Tp).upload().solve(_cabi.default_options(path=2, eps_rel=2e-5))
    assert int(pg.status[0]) == 0
    assert np.abs(pg.rates[0, :, : inst.T].cpu().numpy() - ref).max() <= 5e-3


@pytest.mark.parametrize("path", [1, 2])
def test_zero_limit_row_is_enforced(require_gpu, path):
    """A constraint row with limit 0 (a de-rated line): the reference would enforce it; the EVSEs behind it must get no
    current while the others charge (the relative violation measure has no meaning for such a row)."""
    import adacharge_b200 as ab
    from adacharge_b200.generators import config_c2, caltech_acn_infrastructure
    from oracle import mpc

    infra = caltech_acn_infrastructure()
    infra["constraint_limits"] = np.array(infra["constraint_limits"], dtype=float)
    infra["constraint_limits"][-1] = 0.0  # second pod
    iface = ab.TestingInterface(config_c2(12, infra=infra))
    S, I = iface.active_sessions(), iface.infrastructure_info()
    spec = [("tou_energy_cost", 1, {}), ("total_energy", 0.3, {}), ("demand_charge", 1 / 30, {})]
    aco = ab.AdaptiveChargingOptimization([ab.ObjectiveComponent(getattr(ab, n), c, k) for n, c, k in spec], iface, solver_options=dict(path=path))
    R = aco.solve(S, I, prev_peak=iface.get_prev_peak())
    pod = np.abs(np.asarray(I.constraint_matrix)[-1]) > 0
    assert pod.any() and R[pod].max() <= 1e-4, R[pod].max()
    assert R[~pod].sum() > 100
    Ro = mpc.solve_mpc(spec, S, I, iface, "SOC", False, None, iface.get_prev_peak())
    f, fo = (mpc.evaluate_objective(X, spec, I, iface, S, iface.get_prev_peak()) for X in (R, Ro))
    assert abs(f - fo) <= 1e-4 * abs(fo), (f, fo, aco.last_info)
