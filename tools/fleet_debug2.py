"""Parameter sweep on the replay instances that do not converge (single-session sites at their previous peak)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
import adacharge_b200 as ab
from adacharge_b200 import engine, _cabi
from adacharge_b200.generators import caltech_acn_infrastructure
from adacharge_b200.replay_fast import FleetReplay

obj = [ab.ObjectiveComponent(ab.tou_energy_cost), ab.ObjectiveComponent(ab.total_energy, 0.3), ab.ObjectiveComponent(ab.demand_charge, 1 / 30)]
rp = FleetReplay(caltech_acn_infrastructure(), obj, n_sites=1024, seed0=1000, Tp=160)
TT = int(sys.argv[1]) if len(sys.argv) > 1 else 45
for t in range(0, TT):
    rp.step(t)
h, idx, s, pos, n_sess = rp._pack(TT)
pbc = engine.PackedBatch.from_arrays(rp.site, h, rp.Tp, rp.N).upload().solve(rp.options)
bad = np.nonzero(pbc.status.cpu().numpy() != 0)[0].tolist()
print("cold-unsolved sites at t =", TT, bad)
sites = (bad + [0, 1, 2, 3, 4, 5, 6, 7])[:8]
def sub(h, bs):
    o = {}
    for k, v in h.items():
        o[k] = v if k in ("min_rates", "max_rates") else v[bs]
    return o
hs = sub(h, sites)
np.savez("gpurun_out/fleet_hard2.npz", **hs)
sys.exit(0) if len(sys.argv) > 2 else None
b0 = sites[0]
print("site", b0, "T", h["T"][b0], "E", h["sess_energy"][b0, 0], "len", h["sess_len"][b0, 0], "p0", h["peak_p0"][b0], "w", h["peak_w"][b0],
      "beta", h["beta"][b0, :4], h["beta"][b0, h["T"][b0] - 2 : h["T"][b0]], "E/len*k", h["sess_energy"][b0, 0] / h["sess_len"][b0, 0] * 0.208)
def run(**kw):
    opt = _cabi.default_options(**kw)
    pb = engine.PackedBatch.from_arrays(rp.site, hs, rp.Tp, rp.N).upload().solve(opt)
    it, st, sx = pb.iters.cpu().numpy(), pb.status.cpu().numpy(), pb.stats.cpu().numpy()
    print(kw, "iters", it.tolist(), "st", st.tolist(), "gap", [f"{g:.1e}" for g in sx[:, 2]], "rst", sx[:, 6].astype(int).tolist())
    return pb
run()
for r in (0.005, 0.02, 0.07, 0.2, 0.7, 2.0):
    run(rho0=r, max_rescues=0)
run(max_rescues=0, restart=0)
run(max_rescues=0, alpha=1.0)
run(max_rescues=0, kappa=0.0)
run(max_rescues=0, kappa=1.5)
run(max_rescues=0, check_every=50)
run(max_rescues=0, avg_every=1)
pb = run(max_rescues=0, max_iter=3000)
r = pb.rates[0].cpu().numpy()
i = h["sess_row"][b0, 0]
print("rates row", np.round(r[i, : h["T"][b0]], 3))
