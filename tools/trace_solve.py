"""DEV: where one iteration of the on-chip solve kernel spends its time.  Needs the trace build
(`make -C adacharge_b200/csrc trace` -> tools/build/libadacharge_b200_trace.so): block 0 stamps clock64() per warp at
the loop top (0), after its column-pass work (1), after the barrier (2), after its row / coupling work (3) and after the
second barrier (4) for 16 non-check iterations.  Prints per-warp durations of both passes and the barrier waits."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from adacharge_b200 import _cabi

_cabi.LIB_PATH = os.path.join(ROOT, "tools", "build", "libadacharge_b200_trace.so")
import common
from adacharge_b200 import engine

site, insts, _ = common.build_instances(148, 0)
pb = engine.PackedBatch(site, insts).upload()
opt = _cabi.default_options(max_iter=160, eps_rel=-1.0, eps_abs=-1.0)
for _ in range(2):
    pb.solve(opt)
torch.cuda.synchronize()
L = _cabi.lib()
NIT, NW = 16, 32
buf = (C.c_longlong * (NIT * NW * 8))()
n = L.acb_trace_fetch(buf, NIT * NW * 8)
a = np.frombuffer(buf, dtype=np.int64).reshape(NIT, NW, 8).astype(np.float64)
nw = int((a[0, :, 0] > 0).sum())
full = a[:, :nw, :]
a = a[:, :nw, :5]
MHZ = 1965.0
it0 = a[:, :, 0].min(axis=1, keepdims=True)
rel = (a - it0[:, :, None]) / MHZ  # us since the first warp entered the iteration
IT0 = 111
d = np.diff(a[:, 0, 0]) / MHZ
print(f"warps {nw}; iteration time (loop top to loop top, warp 0), iterations {IT0}..{IT0 + NIT - 2}: " + " ".join(f"{x:.1f}" for x in d) + f" us (check iteration: {125})")
col = (a[:, :, 1] - a[:, :, 0]) / MHZ
w1 = (a[:, :, 2] - a[:, :, 1]) / MHZ
row = (a[:, :, 3] - a[:, :, 2]) / MHZ
w2 = (a[:, :, 4] - a[:, :, 3]) / MHZ
# iteration 4 of the window (it = 105) accumulates the running average: report it separately
for name, sel in (("plain iterations", [i for i in range(NIT) if (IT0 + i) % 5 != 0]), ("averaging iterations (it % 5 == 0, not 125)", [i for i in range(NIT) if (IT0 + i) % 5 == 0 and IT0 + i != 125])):
    print(f"--- {name}: mean over {len(sel)} iterations, per warp [us]")
    print("warp   column   wait1    row/cpl  wait2")
    for w in range(nw):
        print(f"{w:4d} {col[sel, w].mean():8.2f} {w1[sel, w].mean():8.2f} {row[sel, w].mean():8.2f} {w2[sel, w].mean():8.2f}")
    print(f" max  {col[sel].max(axis=1).mean():8.2f} {'':8s} {row[sel].max(axis=1).mean():8.2f}")
    print(f"phase ends (us after iteration start): column done {rel[sel, :, 1].max(axis=1).mean():.2f}, barrier1 released {rel[sel, :, 2].min(axis=1).mean():.2f}, "
          f"row done {rel[sel, :, 3].max(axis=1).mean():.2f}, barrier2 released {rel[sel, :, 4].min(axis=1).mean():.2f}")

# the check iteration (it = 125): stamps 3..7 = row pass done, current candidate evaluated (eval_columns + barrier), Lagrangian
# bound + averaged candidate rows done, averaged candidate evaluated, end of the check path (decision, restart, rebuilt inputs)
ic = 125 - IT0
c = full[ic]
t0 = c[:, 0].min()
names = ["loop top", "column pass", "barrier 1", "general row pass", "eval current + barrier", "Lagrangian + averaged rows", "eval averaged", "decision + rebuild"]
print("--- check iteration 125: per-phase end, max over warps [us after iteration start] and the phase's longest warp [us]")
for k in range(1, 8):
    print(f"{names[k]:28s} end {((c[:, k].max() - t0) / MHZ):7.2f}   longest {((c[:, k] - c[:, k - 1]).max() / MHZ):7.2f}   mean {((c[:, k] - c[:, k - 1]).mean() / MHZ):7.2f}")
