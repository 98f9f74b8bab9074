"""DEV: standard on-chip kernel (path 0) against the experimental compact-bounds kernel (path 3) on the bench workload."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, bench
from adacharge_b200 import _cabi, engine
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2368
site, insts, _ = bench.build_instances(B, 0)
pb = engine.PackedBatch(site, insts).upload()
for path in (0, 3, 0, 3):
    opt = _cabi.default_options(path=path)
    pb.solve(opt); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); pb.solve(opt); e1.record(); torch.cuda.synchronize()
    it, st = pb.iters.cpu().numpy(), pb.status.cpu().numpy()
    print(f"path {path}: {e0.elapsed_time(e1):7.1f} ms for {B} instances; iters mean {it.mean():6.1f} max {it.max():5d}; unsolved {(st != 0).sum()}", flush=True)
