"""DEV: the rate polish's distance estimate against the true distance to the oracle schedule, along one solve:
the same instance is solved with iteration caps 100, 200, ... (rate_tol so small that the polish never stops it) and
the returned estimate `rate_est` is printed next to max |R - R_oracle| (tests/golden/mpc_oracle_golden.json)."""
import json
import os
import sys
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

import adacharge_b200 as ab
from adacharge_b200.generators import config_c1

warnings.simplefilter("ignore")
gold = [g for g in json.load(open(os.path.join(ROOT, "tests", "golden", "mpc_oracle_golden.json"))) if g["config"].startswith("c1")]
caps = [int(a) for a in sys.argv[1:]] or list(range(100, 2001, 100)) + [3000, 5000]
for g in gold:
    iface = ab.TestingInterface(config_c1(g["seed"]))
    S, I = iface.active_sessions(), iface.infrastructure_info()
    obj = [ab.ObjectiveComponent(getattr(ab, n), c, k) for n, c, k in g["objective"]]
    print(f"== {g['config']} seed {g['seed']} equal_share coefficient {g['objective'][1][1]}")
    for mi in caps:
        aco = ab.AdaptiveChargingOptimization(obj, iface, solver_options=dict(max_iter=mi, rate_tol=1e-12, accept_inaccurate=dict(gap=1e9, violation=1e9)))
        R = aco.solve(S, I)
        i = aco.last_info
        print(f"  cap {mi:6d}: iters {i['iters']:6d} status {i['status']} gap {i['gap']:9.2e} restarts {i['restarts']:3d} rate_est {i['rate_est']:9.2e}  true max|dR| {np.abs(R - np.array(g['rates'])).max():9.2e}")
