"""Regenerates tests/golden/replay_degenerate_instances.npz (needs a GPU: the instances are states of the
closed-loop replay).  Runs the 1024-site fleet replay of tools/replay_c4.py to a mid-morning step and keeps the
single-EV sites that sit exactly at their previous peak: remaining energy within 0.1 % of what the sunk peak
lets the EV draw over its remaining stay.  These are the degenerate LPs (|P| << |objective terms|, optimum a few mA
from the flat schedule) that tests/test_gpu_degenerate.py checks against exact LP optima.

    python tests/golden/make_replay_degenerate.py [step=48] [max_instances=8]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import adacharge_b200 as ab  # noqa: E402
from adacharge_b200.generators import caltech_acn_infrastructure  # noqa: E402
from adacharge_b200.replay_fast import FleetReplay  # noqa: E402

step = int(sys.argv[1]) if len(sys.argv) > 1 else 48
keep_n = int(sys.argv[2]) if len(sys.argv) > 2 else 8
obj = [ab.ObjectiveComponent(ab.tou_energy_cost), ab.ObjectiveComponent(ab.total_energy, 0.3), ab.ObjectiveComponent(ab.demand_charge, 1 / 30)]
rp = FleetReplay(caltech_acn_infrastructure(), obj, n_sites=1024, seed0=1000, Tp=160)
for t in range(step):
    rp.step(t)
h, idx, s, pos, n_sess = rp._pack(step)
k = np.asarray(rp.volt) / 1e3
sel = []
for b in np.nonzero(n_sess == 1)[0]:
    i, E, ln, p0 = int(h["sess_row"][b, 0]), float(h["sess_energy"][b, 0]), int(h["sess_len"][b, 0]), float(h["peak_p0"][b])
    if p0 > 0 and abs(E - ln * p0 / k[i]) <= 1e-3 * E:
        sel.append(int(b))
sel = sel[:keep_n]
print(f"step {step}: {len(sel)} degenerate single-EV sites kept: {sel}")
out = {name: v[sel] for name, v in h.items() if name not in ("min_rates", "max_rates")}
mins, maxs = [], []
off = np.zeros_like(out["sess_rate_off"])
for j, b in enumerate(sel):
    p = -(int(h["sess_rate_off"][b, 0]) + 1)
    mins.append(h["min_rates"][p]); maxs.append(h["max_rates"][p])
    off[j, 0] = -(j + 1)
out["sess_rate_off"], out["min_rates"], out["max_rates"] = off, np.array(mins, np.float32), np.array(maxs, np.float32)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "replay_degenerate_instances.npz"), **out)
