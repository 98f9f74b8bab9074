"""DEV: how fast the schedule itself converges on strictly concave objectives (quick_charge + c * equal_share) and what the
rate polish (acb_options.rate_tol) buys: max |R - R_oracle| against the oracle schedules stored in
tests/golden/mpc_oracle_golden.json."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

import adacharge_b200 as ab
from adacharge_b200.generators import config_c1

gold = [g for g in json.load(open(os.path.join(ROOT, "tests", "golden", "mpc_oracle_golden.json"))) if g["config"].startswith("c1")]


def run(g, **opts):
    iface = ab.TestingInterface(config_c1(g["seed"]))
    S, I = iface.active_sessions(), iface.infrastructure_info()
    obj = [ab.ObjectiveComponent(getattr(ab, n), c, k) for n, c, k in g["objective"]]
    aco = ab.AdaptiveChargingOptimization(obj, iface, solver_options=opts)
    try:
        R = aco.solve(S, I)
    except ab.InfeasibilityException as e:
        return None, aco.last_info
    return np.abs(R - np.array(g["rates"])).max(), aco.last_info


for g in gold:
    c = g["objective"][1][1]
    print(f"== {g['config']} seed {g['seed']} equal_share coefficient {c}")
    for name, o in (("class default (eps_rel 2e-5, polish on)", {}), ("eps_rel 1e-4, polish on", dict(eps_rel=1e-4)), ("eps_rel 1e-4, polish off", dict(eps_rel=1e-4, rate_tol=0.0)),
                    ("polish forced (min_qd 0)", dict(eps_rel=1e-4, polish_min_qd=0.0, max_iter=40000)),
                    ("polish forced, rate_tol 1e-4", dict(eps_rel=1e-4, polish_min_qd=0.0, rate_tol=1e-4, max_iter=40000))):
        err, info = run(g, **o)
        print(f"  {name:42s} max|dR| {err if err is None else format(err, '.2e')} A  iters {info['iters']:6d} status {info['status']} gap {info['gap']:.1e} rate_est {info.get('rate_est')}")
    for mi in (100, 200, 400, 800, 1600, 3200, 6400, 12800):
        err, info = run(g, eps_rel=-1.0, eps_abs=-1.0, max_iter=mi, rate_tol=0.0, accept_inaccurate=dict(gap=1e9, violation=1e9))
        print(f"  fixed {mi:6d} iterations: max|dR| {err if err is None else format(err, '.2e')} A  gap {info['gap']:.1e}")
