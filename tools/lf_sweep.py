"""DEV: iteration counts of the on-chip path on load-flattening-dominated Caltech-size instances vs the cold-start penalty."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import adacharge_b200 as ab
from adacharge_b200 import engine, _cabi
from adacharge_b200.generators import config_c2

def batch(objf, B=96):
    insts = []
    for seed in range(B):
        d = config_c2(seed, price_noise=0.2)
        iface = ab.TestingInterface(d)
        S, I = iface.active_sessions(), iface.infrastructure_info()
        T = max(s.arrival_offset + s.remaining_time for s in S)
        aco = ab.AdaptiveChargingOptimization(objf(T), iface)
        insts.append(aco.build_instance(S, I, None, iface.get_prev_peak()))
    return engine.PackedBatch(aco._site_for(I, insts[0]), insts).upload()

ext = lambda T: (40 + 25 * np.sin(np.arange(T) / 30.0)).tolist()
cases = {
    "lf 1.0 + ncp 100": lambda T: [ab.ObjectiveComponent(ab.load_flattening, 1.0, {"external_signal": ext(T)}), ab.ObjectiveComponent(ab.non_completion_penalty, 100.0)],
    "lf 0.01 + tou + energy": lambda T: [ab.ObjectiveComponent(ab.load_flattening, 0.01, {"external_signal": ext(T)}), ab.ObjectiveComponent(ab.tou_energy_cost), ab.ObjectiveComponent(ab.total_energy, 0.3)],
    "lf 1e-4 + qc": lambda T: [ab.ObjectiveComponent(ab.load_flattening, 1e-4, {"external_signal": ext(T)}), ab.ObjectiveComponent(ab.quick_charge)],
    "equal_share 1 + qc": lambda T: [ab.ObjectiveComponent(ab.equal_share, 1.0), ab.ObjectiveComponent(ab.quick_charge)],
    "equal_share 0.01 + tou + energy": lambda T: [ab.ObjectiveComponent(ab.equal_share, 0.01), ab.ObjectiveComponent(ab.tou_energy_cost), ab.ObjectiveComponent(ab.total_energy, 0.3)],
}
for name, objf in cases.items():
    pb = batch(objf)
    for kw in (dict(), dict(rho_curv=0.0), dict(rho_curv=0.3), dict(rho_curv=3.0), dict(rho_curv=0.0, rho0=0.3), dict(rho_curv=0.0, rho0=1.0), dict(rho_curv=0.0, rho0=0.02)):
        pb.solve(_cabi.default_options(**kw)); torch.cuda.synchronize()
        it, st, sx = pb.iters.cpu().numpy(), pb.status.cpu().numpy(), pb.stats.cpu().numpy()
        print(f"{name:34s} {str(kw):40s} iters mean {it.mean():7.1f} max {it.max():6d} solved {int((st == 0).sum())}/{len(st)} rho {sx[:,4].mean():.3g}")
