"""DEV: straggler tail of the bench batch.  (1) data for the relaunch order of a phased solve: the relative gap every
instance has after K iterations against the iterations it finally needs; (2) one launch against the phased solve
(acb_options.phase_iters), which must return bit-identical schedules."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import common
from adacharge_b200 import _cabi, engine

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
site, insts, _ = common.build_instances(B, 0)
pb = engine.PackedBatch(site, insts).upload()


def timed(opt, reps=3):
    pb.solve(opt); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); pb.solve(opt); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


base = _cabi.default_options(phase_iters=0)
t0 = timed(base)
it_full = pb.iters.cpu().numpy().copy(); rates_full = pb.rates.cpu().numpy().copy(); st_full = pb.status.cpu().numpy().copy()
print(f"one launch: {t0:.2f} ms, iters mean {it_full.mean():.1f} max {it_full.max()}, solved {(st_full == 0).sum()}/{B}; ideal {it_full.sum() / 148:.0f} iteration-slots")
out = dict(iters=it_full)
for K in (50, 100, 150, 200):
    timed(_cabi.default_options(phase_iters=0, max_iter=K), reps=1)
    out[f"gap{K}"] = pb.stats[:, 2].cpu().numpy().copy(); out[f"viol{K}"] = pb.stats[:, 3].cpu().numpy().copy(); out[f"it{K}"] = pb.iters.cpu().numpy().copy()
np.savez(os.path.join(ROOT, "gpurun_out", "tail_study2.npz"), **out)
for K in (50, 75, 100, 150, 200, 300):
    t = timed(_cabi.default_options(phase_iters=K))
    it = pb.iters.cpu().numpy(); st = pb.status.cpu().numpy()
    same = np.array_equal(pb.rates.cpu().numpy(), rates_full) and np.array_equal(it, it_full) and np.array_equal(st, st_full)
    print(f"phased, first launch {K:4d} iterations: {t:.2f} ms ({t0 / t:.3f}x), identical to one launch: {same}")
