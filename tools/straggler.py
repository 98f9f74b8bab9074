"""DEV: gap trajectory of the slowest instance of a bench batch."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, common
from adacharge_b200 import _cabi, engine
B = 2368
site, insts, _ = common.build_instances(B, 0)
pb = engine.PackedBatch(site, insts).upload().solve(); torch.cuda.synchronize()
it = pb.iters.cpu().numpy()
order = np.argsort(-it)[:3]
print("slowest:", order, it[order], "sessions", [len(insts[i].sess_row) for i in order])
for i in order[:2]:
    one = engine.PackedBatch(site, [insts[i]]).upload()
    for mi in (250, 500, 1000, 2000, 3000):
        one.solve(_cabi.default_options(max_iter=mi)); torch.cuda.synchronize()
        st = one.stats[0].cpu().numpy()
        print(i, "max_iter", mi, "iters", int(one.iters[0]), "status", int(one.status[0]), "gap %.2e viol %.1e restarts %d avg %d rp %.1e rd %.1e" % (st[2], st[3], st[6], st[7], st[0], st[1]))
    for kw in (dict(rho0=0.2), dict(rho0=0.03), dict(kappa=0.3), dict(kappa=2.0), dict(alpha=1.8), dict(restart=0), dict(avg_every=2), dict(check_every=10)):
        one.solve(_cabi.default_options(**kw)); torch.cuda.synchronize()
        print(i, kw, "iters", int(one.iters[0]), "status", int(one.status[0]))
