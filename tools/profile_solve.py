"""Profiling target: one short launch of the solve kernel (148 C2 instances, iteration cap)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import common
from adacharge_b200 import _cabi, engine

B = int(sys.argv[1]) if len(sys.argv) > 1 else 148
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 300
site, insts, _ = common.build_instances(B, 0)
pb = engine.PackedBatch(site, insts).upload()
opt = _cabi.default_options(max_iter=iters)
for _ in range(2):
    pb.solve(opt)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); pb.solve(opt); e1.record(); torch.cuda.synchronize()
it = pb.iters.cpu().numpy()
print(f"B={B} max_iter={iters}: {e0.elapsed_time(e1):.3f} ms, iters mean {it.mean():.0f}; {e0.elapsed_time(e1)*1e3/it.max():.2f} us/iteration (slowest instance)")
