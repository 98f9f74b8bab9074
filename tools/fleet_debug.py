"""Find the first replay steps with unsolved instances; print their stats and re-solve them cold."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
import adacharge_b200 as ab
from adacharge_b200 import engine, _cabi
from adacharge_b200.generators import caltech_acn_infrastructure
from adacharge_b200.replay_fast import FleetReplay

n_sites = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
t1 = int(sys.argv[2]) if len(sys.argv) > 2 else 288
obj = [ab.ObjectiveComponent(ab.tou_energy_cost), ab.ObjectiveComponent(ab.total_energy, 0.3), ab.ObjectiveComponent(ab.demand_charge, 1 / 30)]
rp = FleetReplay(caltech_acn_infrastructure(), obj, n_sites=n_sites, seed0=1000, Tp=160)
found = 0
orig_step = rp.step
for t in range(0, t1):
    h, idx, s, pos, n_sess = rp._pack(t)
    prev = rp._prev
    rp.step(t)
    if rp.stats.unsolved[-1] == 0:
        continue
    # re-solve this step's batch: warm (as in the replay) and cold, and print the failing rows
    pbw = engine.PackedBatch.from_arrays(rp.site, h, rp.Tp, rp.N, want_warm_out=True)
    rp2_prev, rp._prev = rp._prev, prev
    if prev is not None:
        dev = pbw.device
        pbw.warm = rp._shifted_warm(torch.from_numpy(idx).to(dev), torch.from_numpy(s).to(dev), torch.from_numpy(pos).to(dev), pbw)
    rp._prev = rp2_prev
    pbw.upload().solve(rp.options)
    pbc = engine.PackedBatch.from_arrays(rp.site, h, rp.Tp, rp.N).upload().solve(rp.options)
    stw, stc = pbw.status.cpu().numpy(), pbc.status.cpu().numpy()
    itw, itc = pbw.iters.cpu().numpy(), pbc.iters.cpu().numpy()
    sw, sc = pbw.stats.cpu().numpy(), pbc.stats.cpu().numpy()
    bad = np.nonzero(stw != 0)[0]
    print(f"t={t}: {len(bad)} unsolved warm; cold unsolved {int((stc != 0).sum())}")
    for b in bad[:6]:
        print(f"  site {b}: nS={n_sess[b]} T={h['T'][b]} warm it={itw[b]} st={stw[b]} gap={sw[b,2]:.2e} viol={sw[b,3]:.2e} rho={sw[b,4]:.3g} restarts={sw[b,6]:.0f} | "
              f"cold it={itc[b]} st={stc[b]} gap={sc[b,2]:.2e} viol={sc[b,3]:.2e} rho={sc[b,4]:.3g} p0={h['peak_p0'][b]:.2f} E={h['sess_energy'][b,:n_sess[b]].sum():.0f}")
    found += 1
    if found >= 6:
        break
