"""DEV: throughput of the general path on BASELINE config 5 (1000 EVSEs x 288)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import adacharge_b200 as ab
from adacharge_b200 import engine, _cabi
from adacharge_b200.generators import config_c5, hierarchical_three_phase_network

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
infra = hierarchical_three_phase_network(1000)
insts = []
t0 = time.time()
for seed in range(B):
    d = config_c5(seed, infra=infra)
    iface = ab.TestingInterface(d)
    S, I = iface.active_sessions(), iface.infrastructure_info()
    aco = ab.AdaptiveChargingOptimization([ab.ObjectiveComponent(ab.load_flattening, 1.0, {"external_signal": d["external_signal"]}),
                                           ab.ObjectiveComponent(ab.non_completion_penalty, 100.0)], iface)
    insts.append(aco.build_instance(S, I))
print(f"host packing {time.time()-t0:.1f}s")
site = aco._site_for(I, insts[0])
print("site R", site.R, "NG", site.NG)
pb = engine.PackedBatch(site, insts).upload()
# extra arguments: check_every values to try at the default tolerances (at which iteration does the gap certify?)
extra = [dict(check_every=int(a)) for a in sys.argv[2:]]
for kw in [dict(), dict(max_iter=310, check_every=300, eps_rel=0.0, eps_abs=0.0)] + extra:
  opt = _cabi.default_options(**kw)
  pb.solve(opt); torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record(); pb.solve(opt); e1.record(); torch.cuda.synchronize()
  ms = e0.elapsed_time(e1)
  it = pb.iters.cpu().numpy(); st = pb.status.cpu().numpy()
  print(kw, f"C5 B={B}: {ms:.1f} ms -> {B/ms*1e3:.1f} solves/s; iters mean {it.mean():.0f} max {it.max()}; status {np.bincount(st, minlength=4)}; {ms*1e3/it.max():.1f} us per batch-iteration")
