"""Recompute P and D in float64 on the host from the state the kernel exits with, for the stalled replay instances."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
from adacharge_b200 import engine, _cabi
from adacharge_b200.generators import caltech_acn_infrastructure
from adacharge_b200.interface import InfrastructureInfo

h = dict(np.load(sys.argv[1] if len(sys.argv) > 1 else "tests/golden/replay_degenerate_instances.npz"))
infra = caltech_acn_infrastructure()
info = InfrastructureInfo(np.asarray(infra["constraint_matrix"]), np.asarray(infra["constraint_limits"]), np.asarray(infra["phases"]),
                          np.asarray(infra["voltages"], float), infra["constraint_ids"], infra["station_ids"], np.asarray(infra["max_pilot"]), np.asarray(infra["min_pilot"]))
site = engine.get_site(info, "SOC", False, True)
Tp, N = 160, 54
for kw in (dict(max_iter=1500, max_rescues=0, stall_exit=0), dict(max_iter=1500, max_rescues=0, stall_exit=0, dual_refine=2)):
    opt = _cabi.default_options(**kw)
    pb = engine.PackedBatch.from_arrays(site, h, Tp, N, want_warm_out=True).upload().solve(opt)
    st, it, sx = pb.status.cpu().numpy(), pb.iters.cpu().numpy(), pb.stats.cpu().numpy()
    v1, vc, mu, scal = (pb.warm_out[k].cpu().numpy().astype(np.float64) for k in ("v1", "vc", "mu", "scal"))
    rates = pb.rates.cpu().numpy().astype(np.float64)
    R = vc.shape[1]
    k = np.asarray(infra["voltages"], float) / 1e3
    su = np.linalg.norm(k)
    print(kw)
    for j in range(4):
        T = int(h["T"][j]); i = int(h["sess_row"][j, 0]); E = float(h["sess_energy"][j, 0]); ln = int(h["sess_len"][j, 0])
        cs, rho = sx[j, 5], sx[j, 4]; rho1 = opt.kappa * rho
        w, p0 = float(h["peak_w"][j]) * cs, float(h["peak_p0"][j])
        c = cs * (h["alpha"][j, :T].astype(np.float64) + k[i] * h["beta"][j, :T].astype(np.float64))
        lam = rho1 * mu[j, 0]
        pl = scal[j, 1]
        vu = vc[j, R - 1, :T]; a = vu * su; z = np.minimum(a, pl); y = rho * (vu - z / su)
        rc = c + (k[i] / su) * y + lam
        ub = 32.0
        Dv = np.minimum(0, ub * rc).sum() - lam * E + w * max(z.max(), p0) - (y * z / su).sum()
        # best lam for this y
        base = c + (k[i] / su) * y
        best = max((-L * E + np.minimum(0, ub * (base + L)).sum(), L) for L in np.r_[0.0, np.maximum(-base, 0), lam])
        Dbest = best[0] + w * max(z.max(), p0) - (y * z / su).sum()
        r = rates[j, i, :T]
        lin, pk = (c * r).sum(), w * max((k[i] * r).max(), p0)
        P = lin + pk
        zc = np.clip(v1[j, i, :T] - mu[j, 0], 0, ub)
        print(f" inst {j}: it={it[j]} st={st[j]} kernel gap={sx[j,2]:.2e} | host P={P:.6f} (lin {lin:.3f} pk {pk:.3f}) D={Dv:.6f} D(best lam)={Dbest:.6f} "
              f"gap/max(|P|,.05mag)={(P-Dv)/max(abs(P),0.05*(abs(lin)+pk)):.2e} refined={(P-Dbest)/max(abs(P),0.05*(abs(lin)+pk)):.2e} | rc min {rc.min():.2e} max {rc.max():.2e} "
              f"| r range [{r.min():.4f},{r.max():.4f}] cand range [{zc.min():.4f},{zc.max():.4f}] sum r-E {r.sum()-E:.2e} pl-p0 {pl-p0:.2e} other rows max|vc| {np.abs(vc[j,:R-1]).max():.2e}")
