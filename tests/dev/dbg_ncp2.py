import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import adacharge_b200 as ab
from oracle import mpc
from tests.test_gpu_ncp2 import _case
spec = [("tou_energy_cost", 1, {}), ("non_completion_penalty", 0.5, {"norm": 2}), ("demand_charge", 0.02, {})]
iface = _case("three_phase", 2)
S, I = iface.active_sessions(), iface.infrastructure_info()
Ro = mpc.solve_mpc(spec, S, I, iface, "SOC", False, None, iface.get_prev_peak())
fo = mpc.evaluate_objective(Ro, spec, I, iface, S, iface.get_prev_peak())
print("oracle", fo, [mpc.evaluate_objective(Ro, [o], I, iface, S, iface.get_prev_peak()) for o in spec])
for path in (1, 2):
    for mi in (2000,):
        for extra in ({}, {"restart": 0}, {"max_rescues": 0}, {"rho0": 0.21}, {"rho0": 0.6}, {"check_every": 10}):
            obj = [ab.ObjectiveComponent(getattr(ab, n), c, k) for n, c, k in spec]
            aco = ab.AdaptiveChargingOptimization(obj, iface, solver_options=dict(path=path, max_iter=mi, accept_inaccurate=dict(gap=1e9, violation=1e9), **extra))
            try:
                R = aco.solve(S, I, prev_peak=iface.get_prev_peak())
                f = mpc.evaluate_objective(R, spec, I, iface, S, iface.get_prev_peak())
                print(path, mi, extra, "f", f, "rel", (f - fo) / abs(fo), {k: aco.last_info[k] for k in ("status", "iters", "gap", "violation", "rho")})
            except Exception as e:
                print(path, mi, extra, "EXC", e, aco.last_info)
