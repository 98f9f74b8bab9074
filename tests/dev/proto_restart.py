"""DEV-ONLY numpy model: restarted-averaging ADMM with the rigorous Lagrangian gap as
stopping rule (what the CUDA kernel implements)."""
import sys, time
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests/dev')
import numpy as np
import proto_kernel_model as pk
from proto_admm import pack
from oracle import mpc


def solve(P, rho=0.1, kappa=1.0, alpha=1.6, eps=1e-4, viol_tol=1e-5, max_iter=20000, check=25, m=5, beta=0.5, thr=5.0, verbose=False, use_restart=True):
    N, T = P["N"], P["T"]
    has_pl = P["peak_limit"] is not None; has_u = P["Gamma"] > 0 or len(P["peaks"]) > 0
    site = pk.build_site(P, has_pl, has_u); R, ng, grp, C = site["R"], site["ng"], site["grp"], site["C"]; scales = site["scales"]
    lb, ub = P["lb"], P["ub"]
    first = [np.nonzero(grp == g)[0][0] for g in range(ng)]
    cg = P["c"][first]; cs = 1.0 / max(np.abs(cg).max(), 1e-12); cg = cg * cs; qd = P["qd"] * cs; Gamma = P["Gamma"] * cs
    pk_w = sum(w for w, _ in P["peaks"]) * cs; pk_p0 = P["peaks"][0][1] if P["peaks"] else 0.0
    rows = P["rows"]; Ebar = np.array([e / w for (_, _, _, w, e) in rows]); mus = np.zeros(len(rows))
    M, ML = site["M"], site["ML"]; v1 = lb.copy(); vc = np.zeros((R, T)); plevel = pk_p0
    tt = np.arange(T); masks = [(tt >= a) & (tt < b) for (_, a, b, _, _) in rows]
    cfull = cg[grp]; eq = P["equality"]
    rPL = 2 * M + ML; rU = rPL + (1 if has_pl else 0)

    def mu_map(mv):
        o = np.zeros((N, T))
        for s, (i, a, b, w, e) in enumerate(rows): o[i, a:b] = mv[s]
        return o

    def prox_B(v, mu0):
        mu = mu0.copy()
        for s, (i, a, b, w, e) in enumerate(rows):
            mu[s], _ = pk.newton_prox(v[i], lb[i], ub[i], masks[s], Ebar[s], mu0[s], eq)
        return np.clip(v - mu_map(mu), lb, ub), mu

    def proj_c(v, rho, pl):
        z = v.copy(); r = 0
        for j in range(M):
            lim = P["limits"][j] / scales[r]; nrm = np.hypot(v[r], v[r + 1]); f = np.minimum(1.0, lim / np.maximum(nrm, 1e-30)); z[r] = v[r] * f; z[r + 1] = v[r + 1] * f; r += 2
        for j in range(ML):
            z[r] = np.minimum(v[r], P["limits"][j] / scales[r]); r += 1
        if has_pl:
            z[r] = np.minimum(v[r], P["peak_limit"] / scales[r]); r += 1
        if has_u:
            su = scales[r]; rp = rho / su**2; cur = rp + 2 * Gamma; a = (rp * (v[r] * su) - 2 * Gamma * P["ebar"]) / cur
            z[r] = np.minimum(a, pl) / su if pk_w > 0 else a / su
        return z

    def u_level(v, rho):
        r = rU; su = scales[r]; rp = rho / su**2; cur = rp + 2 * Gamma; a = (rp * (v[r] * su) - 2 * Gamma * P["ebar"]) / cur
        if pk_w <= 0 or a.max() <= pk_p0: return max(a.max(), pk_p0)
        if cur * np.maximum(a - pk_p0, 0).sum() <= pk_w: return pk_p0
        lo, hi = pk_p0, a.max()
        for _ in range(60):
            p = 0.5 * (lo + hi); F = cur * np.maximum(a - p, 0).sum() - pk_w
            if F > 0: lo = p
            else: hi = p
        return p

    def gu(u): return Gamma * ((u + P["ebar"])**2).sum() + pk_w * max(u.max(), pk_p0)

    def primal(z):
        sz = np.zeros((ng, T)); np.add.at(sz, grp, z); Kz = C @ sz
        u = Kz[rU] * scales[rU] if has_u else np.zeros(T)
        Pv = (cfull * z).sum() + qd * (z**2).sum() + (gu(u) if has_u else 0)
        viol = -1.0; r = 0
        for j in range(M): viol = max(viol, (np.hypot(Kz[r], Kz[r + 1]) * scales[r] / P["limits"][j] - 1).max()); r += 2
        for j in range(ML): viol = max(viol, (Kz[r] * scales[r] / P["limits"][j] - 1).max()); r += 1
        if has_pl: viol = max(viol, ((Kz[r] * scales[r] - P["peak_limit"]) / P["peak_limit"]).max())
        return Pv, viol

    def dual_bound(yc, zc, lam):
        rt = cfull + ((C.T @ yc)[grp] if R else 0) + mu_map(lam)
        if qd > 0:
            xs = np.clip(-rt / (2 * qd), lb, ub); phi = qd * xs * xs + rt * xs
        else:
            phi = np.minimum(lb * rt, ub * rt)
        D = phi.sum() - (lam * Ebar).sum(); r = 0
        for j in range(M): D -= (P["limits"][j] / scales[r]) * np.hypot(yc[r], yc[r + 1]).sum(); r += 2
        for j in range(ML): D -= (P["limits"][j] / scales[r]) * yc[r].sum(); r += 1
        if has_pl: D -= (P["peak_limit"] / scales[r] * yc[r]).sum(); r += 1
        if has_u:
            zu = zc[r] * scales[r]; D += gu(zu) - (yc[r] * zc[r]).sum()
        return D

    rho1 = kappa * rho; d = 2 * qd + rho1; Mf = pk.make_M(site, d, rho)
    z1 = np.clip(v1 - mu_map(mus), lb, ub)
    vsum = np.zeros((N, T)); vcsum = np.zeros((R, T)); nsum = 0; Dbest = -np.inf; gap_restart = np.inf; nrestart = 0
    hist = []
    for it in range(1, max_iter + 1):
        qv = 2 * z1 - v1; sq = np.zeros((ng, T)); np.add.at(sq, grp, qv)
        zc = proj_c(vc, rho, plevel); g = rho * (2 * zc - vc)
        ins = np.vstack([rho1 * sq - site["ngrp"][:, None] * cg, g]); outs = Mf @ ins; hgp = outs[:ng] - cg; Kx = outs[ng:] / rho
        x = (rho1 * qv + hgp[grp]) / d; v1n = v1 + alpha * (x - z1)
        z1n, mus = prox_B(v1n, mus)
        vcn = vc + alpha * (Kx - zc)
        if has_u: plevel = u_level(vcn, rho)
        zcn = proj_c(vcn, rho, plevel)
        if it % m == 0:
            vsum += v1n; vcsum += vcn; nsum += 1
        if it % check == 0 or it == max_iter:
            yc = rho * (vcn - zcn)
            Dbest = max(Dbest, dual_bound(yc, zcn, rho1 * mus))
            Pc, vic = primal(z1n)
            za, mua = prox_B(vsum / nsum, mus)
            Pa, via = primal(za)
            sc = lambda Pv: max(abs(Pv), abs(Dbest), 1e-12)
            gc, ga = (Pc - Dbest) / sc(Pc), (Pa - Dbest) / sc(Pa)
            hist.append((it, gc, vic, ga, via, rho, nrestart))
            if verbose: print("%5d cur gap %.2e viol %.1e | avg gap %.2e viol %.1e rho %.3g restarts %d" % hist[-1])
            if gc <= eps and vic <= viol_tol: return z1n, it, hist
            if ga <= eps and via <= viol_tol: return za, it, hist
            rp = max(np.abs(x - z1n).max(), np.abs(Kx - zcn).max() if R else 0)
            d1 = rho1 * ((alpha - 1) * (x - z1) + (z1 - z1n)); dc = rho * ((alpha - 1) * (Kx - zc) + (zc - zcn))
            rd = np.abs(d1).max() + (np.abs(C.T @ dc).max() if R else 0)
            pn = max(np.abs(x).max(), 1e-9); dn = max(1.0, np.abs(rho1 * (v1n - z1n)).max())
            restarted = False
            gabs = Pa - Dbest
            if use_restart and nsum >= 2 and (gabs <= beta * gap_restart) and via <= max(vic, viol_tol) * 1.0 + 1e-3:
                gap_restart = gabs; nrestart += 1; restarted = True
                v1n = vsum / nsum; vcn = vcsum / nsum; z1n, mus = prox_B(v1n, mus)
                if has_u: plevel = u_level(vcn, rho)
                zcn = proj_c(vcn, rho, plevel)
                vsum[:] = 0; vcsum[:] = 0; nsum = 0
            ratio = np.sqrt((rp / pn) / max(rd / dn, 1e-12))
            if (ratio > thr or ratio < 1 / thr):
                rn = float(np.clip(rho * ratio, 1e-4, 1e4)); v1n = z1n + (rho / rn) * (v1n - z1n); vcn = zcn + (rho / rn) * (vcn - zcn)
                rho = rn; rho1 = kappa * rho; d = 2 * qd + rho1; Mf = pk.make_M(site, d, rho)
                vsum[:] = 0; vcsum[:] = 0; nsum = 0
        v1, vc, z1 = v1n, vcn, z1n
    return z1, it, hist


if __name__ == "__main__":
    from adacharge_b200.interface import TestingInterface
    from adacharge_b200.generators import *
    obj2 = [("tou_energy_cost", 1, {}), ("total_energy", 0.3, {}), ("demand_charge", 1 / 30, {})]
    cases = []
    for seed in (0, 1, 2, 3): cases.append((f"c2 noise seed{seed}", config_c2(seed, price_noise=0.2), obj2, {}))
    cases.append(("c2 cap40 seed2", config_c2(2, infra=caltech_acn_infrastructure(transformer_cap=40)), obj2, {}))
    cases.append(("c2 cap30 noise seed5", config_c2(5, infra=caltech_acn_infrastructure(transformer_cap=30), price_noise=0.2), obj2, {}))
    cases.append(("c1 seed0", config_c1(0), [("quick_charge", 1, {}), ("equal_share", 1e-3, {})], {}))
    cases.append(("c1 qc only", config_c1(1), [("quick_charge", 1, {})], {}))
    cases.append(("c2 peak250 lf", config_c2(6, infra=caltech_acn_infrastructure(transformer_cap=60)), [("quick_charge", 1e-3, {}), ("total_energy", 0.3, {}), ("tou_energy_cost", 1, {}), ("load_flattening", 1e-4, {})], dict(peak_limit=250.0)))
    for name, d, obj, kw in cases:
        iface = TestingInterface(d); S = iface.active_sessions(); I = iface.infrastructure_info(); pp = iface.get_prev_peak()
        P = pack(obj, S, I, iface, prev_peak=pp, **kw)
        Ro = mpc.solve_mpc(obj, S, I, iface, prev_peak=pp, **kw); fo = mpc.evaluate_objective(Ro, obj, I, iface, S, pp)
        for ur in (True, False):
            t = time.time(); z, it, hist = solve(P, use_restart=ur, max_iter=6000)
            f = mpc.evaluate_objective(z, obj, I, iface, S, pp); v = mpc.violations(z, S, I, iface, "SOC", kw.get("peak_limit"))
            print(f"{name:24s} restart={ur}: iters {it:5d} true rel err {abs(f-fo)/abs(fo):.2e} viol {v['infrastructure_rel']:.1e} energy {v['energy']:.1e} restarts {hist[-1][-1]} ({time.time()-t:.0f}s)")
