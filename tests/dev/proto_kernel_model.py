"""DEV-ONLY numpy model of the CUDA solve kernel, phase by phase (same state, same
algebra, float32 optional).  Not imported by the product, the tests or the bench."""
import sys, time
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests/dev')
import numpy as np
from oracle import mpc
from proto_admm import pack


def build_site(P, has_pl, has_u):
    """Group-factor the coupling matrix: Khat = C @ Gsel, rows scaled."""
    N = P["N"]; k = P["k"]
    M = len(P["soc"]); ML = len(P["lin_rows"])
    rows, scales, kinds = [], [], []
    for j in range(M):
        s = np.sqrt((P["soc"][j] ** 2).sum() / 2) or 1.0
        rows += [P["soc"][j, 0] / s, P["soc"][j, 1] / s]; scales += [s, s]
    for j in range(ML):
        s = np.linalg.norm(P["lin_rows"][j]) or 1.0
        rows.append(P["lin_rows"][j] / s); scales.append(s)
    if has_pl:
        s = np.sqrt(N); rows.append(np.ones(N) / s); scales.append(s)
    if has_u:
        s = np.linalg.norm(k); rows.append(k / s); scales.append(s)
    K = np.array(rows).reshape(-1, N)
    # groups = distinct columns of [K; k]
    key = np.vstack([K, k[None]]).T
    _, first, grp = np.unique(np.round(key, 12), axis=0, return_index=True, return_inverse=True)
    ng = len(first)
    C = K[:, first]
    lam, U = np.linalg.eigh(K @ K.T) if len(K) else (np.zeros(0), np.zeros((0, 0)))
    return dict(K=K, C=C, grp=grp, ng=ng, ngrp=np.bincount(grp, minlength=ng), kg=k[first], U=U, lam=lam,
                scales=np.array(scales), M=M, ML=ML, has_pl=has_pl, has_u=has_u, R=len(K))


def make_M(site, d, rho):
    U, lam, C = site["U"], site["lam"], site["C"]
    Sinv = U @ np.diag(1.0 / (d / rho + lam)) @ U.T
    X = Sinv @ C
    Mhs = -C.T @ X
    Mhg = (d / rho) * X.T
    Mvs = X
    Mvg = np.eye(site["R"]) - (d / rho) * Sinv
    return np.block([[Mhs, Mhg], [Mvs, Mvg]])


def newton_prox(v, lb, ub, mask, Ebar, mu, equality, max_steps=20, tol=1e-6):
    """Find mu with sum_mask clip(v-mu, lb, ub) = Ebar (or <= with mu >= 0). Returns mu, evals."""
    lo, hi = -np.inf, np.inf
    ev = 0
    if not equality:
        z = np.clip(v, lb, ub); E = z[mask].sum(); ev += 1
        if E <= Ebar * (1 + tol) + 1e-9:
            return 0.0, ev
        lo = 0.0
        if mu <= 0:
            # first guess from the free count at mu = 0
            nf = ((v > lb) & (v < ub) & mask).sum()
            mu = (E - Ebar) / max(nf, 1)
    for _ in range(max_steps):
        z = np.clip(v - mu, lb, ub); E = z[mask].sum(); ev += 1
        r = E - Ebar
        if abs(r) <= tol * max(Ebar, 1.0):
            break
        if r > 0: lo = mu
        else: hi = mu
        nf = ((v - mu > lb) & (v - mu < ub) & mask).sum()
        if nf > 0:
            mun = mu + r / nf
        else:
            mun = np.nan
        if not np.isfinite(mun) or mun <= lo or mun >= hi:
            if np.isfinite(lo) and np.isfinite(hi):
                mun = 0.5 * (lo + hi)
            elif np.isfinite(lo):
                mun = lo + max(1.0, 2 * abs(mu - lo))
            else:
                mun = hi - max(1.0, 2 * abs(hi - mu))
        mu = mun
    return mu, ev


def solve(P, rho=0.1, kappa=1.0, alpha=1.6, eps=1e-4, max_iter=5000, check=25, f32=False, verbose=False):
    ft = np.float32 if f32 else np.float64
    N, T = P["N"], P["T"]
    has_pl = P["peak_limit"] is not None
    has_u = P["Gamma"] > 0 or len(P["peaks"]) > 0
    site = build_site(P, has_pl, has_u)
    R, ng, grp, C = site["R"], site["ng"], site["grp"], site["C"].astype(ft)
    scales = site["scales"]
    lb, ub = P["lb"].astype(ft), P["ub"].astype(ft)
    # group cost cg[g][t]; P["c"] is (N,T) = alpha_t + k_i beta_t so take a representative row
    first = [np.nonzero(grp == g)[0][0] for g in range(ng)]
    cg = P["c"][first].astype(ft)
    cs = ft(1.0 / max(np.abs(cg).max(), 1e-12))
    cg = cg * cs; qd = ft(P["qd"] * cs); Gamma = P["Gamma"] * cs
    pk_w = sum(w for w, _ in P["peaks"]) * cs
    pk_p0 = P["peaks"][0][1] if P["peaks"] else 0.0
    rows = P["rows"]
    Ebar = np.array([e / w for (_, _, _, w, e) in rows])
    mus = np.zeros(len(rows))
    lims = []
    r = 0
    M, ML = site["M"], site["ML"]
    v1 = lb.copy()
    vc = np.zeros((R, T), ft)
    plevel = 0.0
    tt = np.arange(T)
    masks = [(tt >= a) & (tt < b) for (_, a, b, _, _) in rows]

    def mu_map():
        m = np.zeros((N, T), ft)
        for s, (i, a, b, w, e) in enumerate(rows):
            m[i, a:b] = mus[s]
        return m

    def proj_c(v, rho, plevel_in):
        z = v.copy(); r = 0
        for j in range(M):
            lim = P["limits"][j] / scales[r]
            nrm = np.hypot(v[r], v[r + 1]); f = np.minimum(1.0, lim / np.maximum(nrm, 1e-30))
            z[r] = v[r] * f; z[r + 1] = v[r + 1] * f; r += 2
        for j in range(ML):
            z[r] = np.minimum(v[r], P["limits"][j] / scales[r]); r += 1
        if has_pl:
            z[r] = np.minimum(v[r], P["peak_limit"] / scales[r]); r += 1
        if has_u:
            su = scales[r]; rp = rho / su**2; cur = rp + 2 * Gamma
            a = (rp * (v[r] * su) - 2 * Gamma * P["ebar"]) / cur
            z[r] = np.minimum(a, plevel_in) / su if pk_w > 0 else a / su
        return z

    def u_level(v, rho, p):
        """root-find the cap level for the aggregate-power row."""
        r = R - 1; su = scales[r]; rp = rho / su**2; cur = rp + 2 * Gamma
        a = (rp * (v[r] * su) - 2 * Gamma * P["ebar"]) / cur
        if pk_w <= 0 or a.max() <= pk_p0:
            return max(a.max(), pk_p0)
        if cur * np.maximum(a - pk_p0, 0).sum() <= pk_w:
            return pk_p0
        p = min(max(p, pk_p0), a.max())
        lo, hi = pk_p0, a.max()
        for _ in range(30):
            F = cur * np.maximum(a - p, 0).sum() - pk_w
            if abs(F) <= 1e-6 * pk_w: break
            if F > 0: lo = p
            else: hi = p
            na = (a > p).sum()
            pn = p + F / (cur * na) if na > 0 else np.nan
            if not np.isfinite(pn) or pn <= lo or pn >= hi: pn = 0.5 * (lo + hi)
            p = pn
        return p

    rho1 = kappa * rho; d = 2 * qd + rho1
    Mf = make_M(site, d, rho).astype(ft)
    hist = []; evals = 0
    z1 = np.clip(v1 - mu_map(), lb, ub)
    for it in range(1, max_iter + 1):
        # S1: q, partial sums
        qv = 2 * z1 - v1
        sq = np.zeros((ng, T), ft); np.add.at(sq, grp, qv)
        zc = proj_c(vc, rho, plevel)
        g = rho * (2 * zc - vc)
        # S2: column pass
        ins = np.vstack([rho1 * sq - site["ngrp"][:, None] * cg, g])
        outs = Mf @ ins
        hgp = outs[:ng] - cg
        Kx = outs[ng:] / rho
        # S3
        x = (rho1 * qv + hgp[grp]) / d
        v1n = v1 + alpha * (x - z1)
        musn = mus.copy()
        for s, (i, a, b, w, e) in enumerate(rows):
            musn[s], ev = newton_prox(v1n[i], lb[i], ub[i], masks[s], Ebar[s], mus[s], P["equality"]); evals += ev
        mus_old = mus; mus = musn
        z1n = np.clip(v1n - mu_map(), lb, ub)
        vcn = vc + alpha * (Kx - zc)
        if has_u:
            plevel = u_level(vcn, rho, plevel)
        zcn = proj_c(vcn, rho, plevel)
        if it % check == 0 or it == max_iter:
            rp = max(np.abs(x - z1n).max(), np.abs(Kx - zcn).max() if R else 0)
            d1 = rho1 * ((alpha - 1) * (x - z1) + (z1 - z1n))
            dc = rho * ((alpha - 1) * (Kx - zc) + (zc - zcn))
            rd_exact = np.abs(d1 + (C.T @ dc)[grp]).max()
            rd = np.abs(d1).max() + (np.abs(C.T @ dc).max() if R else 0)
            y1 = rho1 * (v1n - z1n); yc = rho * (vcn - zcn)
            pn = max(np.abs(x).max(), np.abs(z1n).max(), 1e-9)
            dn = max(np.abs(cg).max(), np.abs(y1).max(), 1e-9)
            gap = 2 * qd * (x * x).sum() + (cg[grp] * x).sum() + (y1 * z1n).sum() + (yc * zcn).sum()
            pob = qd * (x * x).sum() + (cg[grp] * x).sum()
            gsc = max(abs(pob), 1e-9)
            # violation of z1n
            sz = np.zeros((ng, T), ft); np.add.at(sz, grp, z1n)
            Kz = C @ sz; viol = -1.0; r = 0
            for j in range(M):
                viol = max(viol, (np.hypot(Kz[r], Kz[r + 1]) * scales[r] / P["limits"][j] - 1).max()); r += 2
            for j in range(ML):
                viol = max(viol, (Kz[r] * scales[r] / P["limits"][j] - 1).max()); r += 1
            if has_pl:
                viol = max(viol, ((Kz[r] * scales[r] - P["peak_limit"]) / P["peak_limit"]).max()); r += 1
            hist.append((it, rp / pn, rd / dn, rd_exact / dn, gap / gsc, viol, rho))
            if verbose:
                print(it, f"rp {rp/pn:.2e} rd {rd/dn:.2e} (exact {rd_exact/dn:.2e}) gap {gap/gsc:.2e} viol {viol:.2e} rho {rho:.3g}")
            if rp / pn < eps and rd / dn < eps and abs(gap) / gsc < eps and viol < 1e-5:
                v1, vc, z1 = v1n, vcn, z1n
                break
            ratio = np.sqrt((rp / pn) / max(rd / dn, 1e-12))
            if ratio > 5 or ratio < 0.2:
                rho_new = float(np.clip(rho * ratio, 1e-4, 1e4))
                # rescale stored v so that y is preserved
                v1n = z1n + (rho / rho_new) * (v1n - z1n)
                vcn = zcn + (rho / rho_new) * (vcn - zcn)
                rho = rho_new; rho1 = kappa * rho; d = 2 * qd + rho1
                Mf = make_M(site, d, rho).astype(ft)
        v1, vc, z1 = v1n.astype(ft), vcn.astype(ft), z1n.astype(ft)
    return z1, it, hist, evals / max(1, it * len(rows))


if __name__ == "__main__":
    from adacharge_b200.interface import TestingInterface
    from adacharge_b200.generators import *
    which = sys.argv[1]; seed = int(sys.argv[2]); cap = float(sys.argv[3]) if len(sys.argv) > 3 else 150
    f32 = len(sys.argv) > 4 and sys.argv[4] == "f32"
    if which == "c2":
        d = config_c2(seed=seed, infra=caltech_acn_infrastructure(transformer_cap=cap))
        obj = [("tou_energy_cost", 1, {}), ("total_energy", 0.3, {}), ("demand_charge", 1 / 30, {})]
    else:
        d = config_c1(seed=seed); obj = [("quick_charge", 1, {}), ("equal_share", 1e-3, {})]
    iface = TestingInterface(d); S = iface.active_sessions(); I = iface.infrastructure_info(); pp = iface.get_prev_peak()
    P = pack(obj, S, I, iface, prev_peak=pp)
    Ro = mpc.solve_mpc(obj, S, I, iface, prev_peak=pp); fo = mpc.evaluate_objective(Ro, obj, I, iface)
    t = time.time()
    z, it, hist, ev = solve(P, f32=f32, verbose=True)
    f = mpc.evaluate_objective(z.astype(float), obj, I, iface)
    print(f"iters {it} relobj {abs(f-fo)/abs(fo):.2e} viol {mpc.violations(z.astype(float),S,I,iface)} maxdiff {np.abs(z-Ro).max():.4f} evals/row/iter {ev:.2f} {time.time()-t:.1f}s")
