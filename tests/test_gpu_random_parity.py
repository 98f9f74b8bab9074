"""Seeded random small instances (tests/scenarios.py::random_scenario) through the drop-in class on the
GPU against the float64 oracle: objective within 1e-4 of |f*| (the north-star bar), bounds / energy /
infrastructure / peak rows satisfied.  Only where the objective's terms cancel (|f*| below 1e-3 of the sum of
their magnitudes: a sunk demand charge against the energy revenue) is the error taken relative to the term
scale instead -- 1e-4 of |f*| would then ask for more digits than the terms carry -- and those seeds are listed."""
import numpy as np
import pytest

import adacharge_b200 as ab
from oracle import mpc
from tests.scenarios import make_interface, random_scenario

pytestmark = pytest.mark.gpu
OBJ_TOL, VIOL_TOL = 1e-4, 1e-5


@pytest.mark.parametrize("seed", range(40))
def test_random_instance_matches_oracle(require_gpu, seed):
    sc = random_scenario(seed)
    iface = make_interface(sc)
    S, I = iface.active_sessions(), iface.infrastructure_info()
    ct, eq, pl, pp = sc.get("constraint_type", "SOC"), sc.get("equality", False), sc.get("peak_limit"), iface.get_prev_peak()
    comps = [ab.ObjectiveComponent(getattr(ab, n), c, k) for n, c, k in sc["objective"]]
    aco = ab.AdaptiveChargingOptimization(comps, iface, ct, eq)
    R = aco.solve(S, I, peak_limit=pl, prev_peak=pp)
    assert R.shape == (I.num_stations, mpc.horizon(S))
    v = mpc.violations(R, S, I, iface, ct, pl, eq)
    assert v["lb"] <= 1e-5 and v["ub"] <= 1e-5, v
    assert v["infrastructure_rel"] <= VIOL_TOL and v.get("peak_rel", 0) <= VIOL_TOL, v
    assert v["energy"] <= 1e-4, v  # kWh over the request (the reference's tests: 1e-4, t_aco.py:53-65)
    Ro = mpc.solve_mpc(sc["objective"], S, I, iface, ct, eq, pl, pp)
    f, fo = (mpc.evaluate_objective(X, sc["objective"], I, iface, S, pp) for X in (R, Ro))
    mag = sum(abs(mpc.evaluate_objective(Ro, [o], I, iface, S, pp)) for o in sc["objective"])
    if abs(fo) > 1e-3 * mag:
        assert abs(f - fo) <= OBJ_TOL * abs(fo) + 1e-7, (f, fo, mag, aco.last_info)  # (1e-7: optima that are exactly 0)
    else:
        CANCELLING.append(seed)
        assert abs(f - fo) <= OBJ_TOL * 0.05 * mag + 1e-7, (f, fo, mag, aco.last_info)


CANCELLING = []


def test_cancelling_instances_are_the_exception(require_gpu):
    """(runs after the parametrised cases) at most a few of the 40 seeds fall under the term-scale rule."""
    assert len(CANCELLING) <= 4, CANCELLING
