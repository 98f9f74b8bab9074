import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import adacharge_b200 as ab
from adacharge_b200.generators import caltech_acn_infrastructure
from adacharge_b200.replay import SiteReplay, ReplayStats
obj = [ab.ObjectiveComponent(ab.quick_charge), ab.ObjectiveComponent(ab.equal_share, 1e-6)]
for so in (dict(), dict(stall_checks=0)):
    rp = SiteReplay(caltech_acn_infrastructure(), obj, n_sites=4, steps=288, seed0=7, warm_start=True, solver_options=so)
    st = ReplayStats(); bad = 0
    for t in range(60, 288):
        rp.step(t, st)
        if st.status and (st.status[-1] != 0).any():
            bad += 1
            if bad <= 3: print(so, "t", t, "status", st.status[-1], "iters", st.iters[-1], "\n", rp.last_stats[st.status[-1] != 0])
    print(so, "bad steps", bad, "mean iters", np.concatenate(st.iters).mean())
