"""DEV: iteration counts of the bench batch with host-side features, to study batch ordering against the straggler tail."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, bench
from adacharge_b200 import _cabi, engine
B = 4096
site, insts, _ = bench.build_instances(B, 0)
pb = engine.PackedBatch(site, insts).upload()
pb.solve(); torch.cuda.synchronize()
it = pb.iters.cpu().numpy()
feat = np.array([[len(i.sess_row), i.sess_energy.sum(), i.T, i.peak_p0, (i.sess_len).sum(), i.sess_energy.sum() / max(i.sess_len.sum(), 1), np.abs(i.beta).max()] for i in insts])
np.savez("gpurun_out/tail_study.npz", iters=it, feat=feat)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); pb.solve(); e1.record(); torch.cuda.synchronize(); print("index order ms", e0.elapsed_time(e1))
# oracle ordering: longest first (needs the iteration counts: upper bound on what ordering can give)
order = np.argsort(-it)
insts2 = [insts[i] for i in order]
pb2 = engine.PackedBatch(site, insts2).upload(); pb2.solve(); torch.cuda.synchronize()
e0.record(); pb2.solve(); e1.record(); torch.cuda.synchronize(); print("longest-first (oracle order) ms", e0.elapsed_time(e1))
order = np.argsort(it)
pb3 = engine.PackedBatch(site, [insts[i] for i in order]).upload(); pb3.solve(); torch.cuda.synchronize()
e0.record(); pb3.solve(); e1.record(); torch.cuda.synchronize(); print("shortest-first ms", e0.elapsed_time(e1))
