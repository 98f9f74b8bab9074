"""DEV-ONLY: the replay instances the kernel stalls on (gpurun_out/fleet_hard.npz), in the numpy model."""
import sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests/dev')
import numpy as np
import proto_restart as pr
from oracle import mpc
from adacharge_b200.generators import caltech_acn_infrastructure
from adacharge_b200.interface import InfrastructureInfo

def load(j, path="/root/repo/gpurun_out/fleet_hard.npz"):
    h = np.load(path)
    infra = caltech_acn_infrastructure()
    info = InfrastructureInfo(np.asarray(infra["constraint_matrix"]), np.asarray(infra["constraint_limits"]), np.asarray(infra["phases"]),
                              np.asarray(infra["voltages"], float), infra["constraint_ids"], infra["station_ids"], np.asarray(infra["max_pilot"]), np.asarray(infra["min_pilot"]))
    N = len(infra["station_ids"]); T = int(h["T"][j]); nS = int(h["n_sessions"][j])
    k = np.asarray(infra["voltages"], float) / 1e3
    lb = np.zeros((N, T)); ub = np.zeros((N, T)); rows = []
    for s in range(nS):
        i = int(h["sess_row"][j, s]); a = int(h["sess_start"][j, s]); ln = int(h["sess_len"][j, s]); off = int(h["sess_rate_off"][j, s])
        ub[i, a:a + ln] = h["max_rates"][-(off + 1)]
        w = infra["voltages"][i] * 5 / 1e3 / 60
        rows.append((i, a, a + ln, w, float(h["sess_energy"][j, s]) * w))
    c = h["alpha"][j, :T][None, :].astype(float) + k[:, None] * h["beta"][j, :T][None, :].astype(float)
    P = dict(N=N, T=T, lb=lb, ub=ub, c=c, qd=0.0, k=k, rows=rows, equality=False, Gamma=0.0, ebar=np.zeros(T),
             peaks=[(float(h["peak_w"][j]), float(h["peak_p0"][j]))], soc=mpc.soc_rows(info), lin_rows=np.zeros((0, N)),
             limits=np.asarray(infra["constraint_limits"], float), peak_limit=None)
    return P

if __name__ == "__main__":
    j = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    P = load(j)
    print("T", P["T"], "rows", P["rows"], "peaks", P["peaks"])
    z, it, hist = pr.solve(P, rho=0.07, kappa=0.7, alpha=1.7, thr=1e9, max_iter=int(sys.argv[2]) if len(sys.argv) > 2 else 4000, verbose=False)
    for hh in hist[:: max(1, len(hist) // 40)]:
        print("%5d cur gap %.2e viol %.1e | avg gap %.2e viol %.1e rho %.3g restarts %d" % hh)
    print("iters", it)
