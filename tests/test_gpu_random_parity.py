"""Seeded random small instances (tests/scenarios.py::random_scenario) through the drop-in class on the
GPU against the float64 oracle: objective within 1e-4 (relative to max(|f*|, 5 % of the objective's term
magnitudes): the solver's documented tolerance scale), bounds/energy/infrastructure/peak rows satisfied."""
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
    assert v["energy"] <= 2e-4, v
    Ro = mpc.solve_mpc(sc["objective"], S, I, iface, ct, eq, pl, pp)
    f, fo = (mpc.evaluate_objective(X, sc["objective"], I, iface, S, pp) for X in (R, Ro))
    mag = sum(abs(mpc.evaluate_objective(Ro, [o], I, iface, S, pp)) for o in sc["objective"])
    assert abs(f - fo) <= OBJ_TOL * max(abs(fo), 0.05 * mag) + 1e-7, (f, fo, mag, aco.last_info)
