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

_cabi.LIB_PATH = os.environ.get("ACB_TRACE_LIB", os.path.join(ROOT, "tools", "build", "libadacharge_b200_trace.so"))
IT0 = int(os.environ.get("ACB_TR_IT0", "111"))  # must match the build (make trace TRACE_DEFS=-DACB_TR_IT0=...)
import common
from adacharge_b200 import engine

site, insts, _ = common.build_instances(148, 0)
pb = engine.PackedBatch(site, insts).upload()
opt = _cabi.default_options(max_iter=160, eps_rel=-1.0, eps_abs=-1.0)
for _ in range(2):
    pb.solve(opt)
torch.cuda.synchronize()
L = _cabi.lib()
NIT, NW, NS = 16, 32, 16
buf = (C.c_longlong * (NIT * NW * NS))()
n = L.acb_trace_fetch(buf, NIT * NW * NS)
a = np.frombuffer(buf, dtype=np.int64).reshape(NIT, NW, NS).astype(np.float64)
nw = int((a[0, :, 0] > 0).sum())
raw_flag = a[:, 31, 15].copy()
full = a[:, :nw, :]
a = a[:, :nw, :5]
MHZ = 1965.0
it0 = a[:, :, 0].min(axis=1, keepdims=True)
rel = (a - it0[:, :, None]) / MHZ  # us since the first warp entered the iteration
d = np.diff(a[:, 0, 0]) / MHZ
print(f"warps {nw}; iteration time (loop top to loop top, warp 0), iterations {IT0}..{IT0 + NIT - 2}: " + " ".join(f"{x:.1f}" for x in d) + f" us (check iterations: multiples of 25)")
col = (a[:, :, 1] - a[:, :, 0]) / MHZ
w1 = (a[:, :, 2] - a[:, :, 1]) / MHZ
row = (a[:, :, 3] - a[:, :, 2]) / MHZ
w2 = (a[:, :, 4] - a[:, :, 3]) / MHZ
# iteration 4 of the window (it = 105) accumulates the running average: report it separately
for name, sel in (("plain iterations", [i for i in range(NIT) if (IT0 + i) % 5 != 0]), ("averaging iterations (it % 5 == 0, no check)", [i for i in range(NIT) if (IT0 + i) % 5 == 0 and (IT0 + i) % 25 != 0])):
    print(f"--- {name}: mean over {len(sel)} iterations, per warp [us]")
    print("warp   column   wait1    row/cpl  wait2")
    for w in range(nw):
        print(f"{w:4d} {col[sel, w].mean():8.2f} {w1[sel, w].mean():8.2f} {row[sel, w].mean():8.2f} {w2[sel, w].mean():8.2f}")
    print(f" max  {col[sel].max(axis=1).mean():8.2f} {'':8s} {row[sel].max(axis=1).mean():8.2f}")
    print(f"phase ends (us after iteration start): column done {rel[sel, :, 1].max(axis=1).mean():.2f}, barrier1 released {rel[sel, :, 2].min(axis=1).mean():.2f}, "
          f"row done {rel[sel, :, 3].max(axis=1).mean():.2f}, barrier2 released {rel[sel, :, 4].min(axis=1).mean():.2f}")

# the check iterations of the window: stamps 3.. = row + coupling pass done, current candidate evaluated (eval_columns + barrier),
# Lagrangian bound + averaged candidate rows done, averaged candidate evaluated, reductions + barrier, decision + barrier,
# restart / penalty change, partial sums rebuilt, coupling inputs rebuilt + barrier
names = ["loop top", "column pass", "barrier 1", "general row + coupling pass", "eval current + barrier", "Lagrangian + averaged rows", "eval averaged",
         "reductions + barrier", "decision (thread 0) + barrier", "restart / penalty change", "write_part_q", "write_gin + barrier"]
for ic in range(NIT):
    it = IT0 + ic
    if it % 25 != 0:
        continue
    c = full[ic]
    t0 = c[:, 0].min()
    print(f"--- check iteration {it} (decision flag {int(raw_flag[ic])}): per-phase end, max over warps [us after iteration start]; the phase's longest / mean warp [us]")
    for k in range(1, 12):
        print(f"{names[k]:30s} end {((c[:, k].max() - t0) / MHZ):7.2f}   longest {((c[:, k] - c[:, k - 1]).max() / MHZ):7.2f}   mean {((c[:, k] - c[:, k - 1]).mean() / MHZ):7.2f}")
    # inside the Lagrangian phase, first row of each row warp: stamps 12..14 = bounds and running average loaded, Lagrangian terms done,
    # averaged candidate's multiplier found; the rest of the row (its z, partial sums, objective) runs until the next row starts
    if (c[:, 12] > 0).any():
        w = c[:, 12] > 0
        print(f"  first row of a row warp: loads {((c[w, 12] - c[w, 4]).mean() / MHZ):.2f} us, Lagrangian terms {((c[w, 13] - c[w, 12]).mean() / MHZ):.2f} us, "
              f"averaged multiplier {((c[w, 14] - c[w, 13]).mean() / MHZ):.2f} us; all three rows {((c[w, 5] - c[w, 4]).mean() / MHZ):.2f} us")
