"""DEV: iterations / time to certificate on the bench workload for solver parameter variants.
   python tools/param_sweep.py "dict(newton_rel=0.1)" "dict(check_every=30)" ..."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, common
from adacharge_b200 import _cabi, engine
B = 2368
site, insts, _ = common.build_instances(B, 0)
pb = engine.PackedBatch(site, insts).upload()
def run(**kw):
    opt = _cabi.default_options(**kw)
    pb.solve(opt); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); pb.solve(opt); e1.record(); torch.cuda.synchronize()
    it = pb.iters.cpu().numpy(); st = pb.status.cpu().numpy()
    print(f"{str(kw):70s} {e0.elapsed_time(e1):7.1f} ms  iters mean {it.mean():6.0f} p90 {np.percentile(it,90):6.0f} max {it.max():6d} unsolved {(st!=0).sum()}  us/iter-slot {e0.elapsed_time(e1)*1e3*148/it.sum():.2f}")
run()
variants = [eval(a) for a in sys.argv[1:]] or [dict(check_every=20), dict(check_every=30), dict(check_every=40)]
for kw in variants:
    run(**kw)
