import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import adacharge_b200 as ab
from adacharge_b200.generators import caltech_acn_infrastructure
from adacharge_b200.replay import SiteReplay, ReplayStats
obj = [ab.ObjectiveComponent(ab.tou_energy_cost), ab.ObjectiveComponent(ab.total_energy, 0.3), ab.ObjectiveComponent(ab.demand_charge, 1 / 30)]
for warm in (False, True):
    rp = SiteReplay(caltech_acn_infrastructure(), obj, n_sites=6, steps=288, seed0=100, warm_start=warm, solver_options=dict(max_iter=20000))
    st = ReplayStats()
    for t in range(96, 136):
        rp.step(t, st)
        s, it = st.status[-1], st.iters[-1]
        if (s != 0).any():
            print("warm", warm, "t", t, "status", s, "iters", it)
            print(rp.last_stats[s != 0])
    print("warm", warm, "mean iters", np.concatenate(st.iters).mean())
