"""DEV: cost of the check path = (time with check_every=25) - (time with one check) over 300 iterations."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import common
from adacharge_b200 import _cabi, engine
site, insts, _ = common.build_instances(148, 0)
pb = engine.PackedBatch(site, insts).upload()
def t(max_iter=300, **kw):
    opt = _cabi.default_options(max_iter=max_iter, eps_rel=1e-12, eps_abs=0.0, **kw)
    pb.solve(opt); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); pb.solve(opt); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)
for mi in (25, 50, 100, 200, 400, 800):
    print('max_iter', mi, f"{t(max_iter=mi, check_every=1000, restart=0, adapt_rho=0):.3f} ms")
for kw in (dict(check_every=25), dict(check_every=300), dict(check_every=25, restart=0), dict(check_every=25, adapt_rho=0), dict(check_every=25, restart=0, adapt_rho=0), dict(check_every=5, restart=0, adapt_rho=0)):
    print(kw, f"{t(**kw):.3f} ms")
