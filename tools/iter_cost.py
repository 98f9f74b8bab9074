import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, common
from adacharge_b200 import _cabi, engine
if len(sys.argv) > 2:  # e.g. `nocheck`: tools/build/libadacharge_b200_nocheck.so (make -C adacharge_b200/csrc nocheck)
    _cabi.LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "build", f"libadacharge_b200_{sys.argv[2]}.so")
site, insts, _ = common.build_instances(148, 0)
pb = engine.PackedBatch(site, insts).upload()
def t(mi):
    opt = _cabi.default_options(max_iter=mi, check_every=100000, restart=0, adapt_rho=0, eps_rel=1e-12, eps_abs=0.0)
    pb.solve(opt); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); pb.solve(opt); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)
a, b = t(100), t(500)
print(f"{sys.argv[1] if len(sys.argv)>1 else ''}: {(b-a)/400*1e3:.2f} us per iteration (slowest block)")
