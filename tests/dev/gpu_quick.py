"""DEV: first GPU contact — CUDA solve vs oracle on a few instances."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import adacharge_b200 as ab
from adacharge_b200.generators import *
from adacharge_b200 import engine, _cabi
from oracle import mpc

def run(name, d, obj, equality=False, peak_limit=None, opts=None):
    iface = ab.TestingInterface(d); S = iface.active_sessions(); I = iface.infrastructure_info()
    pp = iface.get_prev_peak()
    aco = ab.AdaptiveChargingOptimization(obj, iface, enforce_energy_equality=equality, solver_options=opts or {})
    torch.cuda.synchronize(); t = time.time()
    try:
        R = aco.solve(S, I, peak_limit=peak_limit, prev_peak=pp)
    except ab.InfeasibilityException as e:
        print(name, "INFEASIBLE", e, aco.last_info); return
    torch.cuda.synchronize(); dt = time.time() - t
    oobj = [(c.function.__name__, c.coefficient, c.kwargs) for c in obj]
    t = time.time(); Ro = mpc.solve_mpc(oobj, S, I, iface, "SOC", equality, peak_limit, pp); to = time.time() - t
    f, fo = mpc.evaluate_objective(R, oobj, I, iface, S, pp), mpc.evaluate_objective(Ro, oobj, I, iface, S, pp)
    print(f"{name}: gpu {dt*1e3:.1f} ms (oracle {to:.1f}s) info {aco.last_info}\n    obj {f:.6f} oracle {fo:.6f} rel {abs(f-fo)/max(1e-12,abs(fo)):.2e} maxdiff {np.abs(R-Ro).max():.4f} viol {mpc.violations(R,S,I,iface,'SOC',peak_limit,equality)}")

OC = ab.ObjectiveComponent
sess = session_generator(2,[0]*2,[12]*2,[3.3]*2,[3.3]*2,[32]*2)
d = {"active_sessions":sess,"infrastructure_info":single_phase_single_constraint(2,64),"current_time":0,"period":5}
run("kat1", d, [OC(ab.quick_charge)])
run("kat1-eq", d, [OC(ab.quick_charge)], equality=True)
run("kat1-peak32", d, [OC(ab.quick_charge)], peak_limit=32)
run("c1", config_c1(0), [OC(ab.quick_charge), OC(ab.equal_share, 1e-3)])
obj2 = [OC(ab.tou_energy_cost), OC(ab.total_energy, 0.3), OC(ab.demand_charge, 1/30)]
run("c2", config_c2(1), obj2)
run("c2-cap40", config_c2(2, infra=caltech_acn_infrastructure(transformer_cap=40)), obj2)
# batch timing
for cfg, obj, B in (("c1", [OC(ab.quick_charge), OC(ab.equal_share, 1e-3)], 592), ("c2", obj2, 296)):
    insts = []
    for seed in range(B):
        d = config_c1(seed) if cfg == "c1" else config_c2(seed, price_noise=0.2)
        iface = ab.TestingInterface(d); S = iface.active_sessions(); I = iface.infrastructure_info()
        aco = ab.AdaptiveChargingOptimization(obj, iface)
        insts.append(aco.build_instance(S, I, None, iface.get_prev_peak()))
    site = aco._site_for(I, insts[0])
    pb = engine.PackedBatch(site, insts).upload()
    pb.solve(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); pb.solve(); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    it = pb.iters.cpu().numpy(); st = pb.status.cpu().numpy()
    print(f"batch {cfg} B={B}: {ms:.2f} ms -> {B/ms*1e3:.0f} solves/s; iters mean {it.mean():.0f} max {it.max()} status counts {np.bincount(st, minlength=4)}; restarts mean {pb.stats[:,6].mean().item():.1f} averaged {pb.stats[:,7].mean().item():.2f}")
