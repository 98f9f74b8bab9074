"""DEV-ONLY numpy model of the device algorithm (used to choose parameters and to debug
the CUDA kernel; not imported by the product, tests or bench)."""
import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np
from oracle import mpc


def pack(objective, sessions, infra, iface, constraint_type="SOC", equality=False, peak_limit=None, prev_peak=0):
    T = mpc.horizon(sessions)
    N = len(infra.station_ids)
    lb, ub = mpc.bounds(sessions, infra.station_ids, T)
    terms = mpc.objective_terms(objective, infra, iface, T, sessions, prev_peak)
    k = np.asarray(infra.voltages, float) / 1e3
    rows = mpc.session_rows(sessions, infra, iface.period)
    P = dict(N=N, T=T, lb=lb, ub=ub, c=terms["lin"], qd=terms["diag_q"], k=k, rows=rows, equality=equality)
    G = sum(g for g, _ in terms["agg"])
    P["Gamma"] = G
    P["ebar"] = sum(g * e for g, e in terms["agg"]) / G if G > 0 else np.zeros(T)
    P["peaks"] = terms["peaks"]
    if mpc.has_infrastructure(infra):
        if constraint_type == "SOC":
            P["soc"] = mpc.soc_rows(infra)  # (M,2,N)
            P["lin_rows"] = np.zeros((0, N))
        else:
            P["soc"] = np.zeros((0, 2, N))
            P["lin_rows"] = np.abs(np.asarray(infra.constraint_matrix, float))
        P["limits"] = np.asarray(infra.constraint_limits, float)
    else:
        P["soc"] = np.zeros((0, 2, N)); P["lin_rows"] = np.zeros((0, N)); P["limits"] = np.zeros(0)
    P["peak_limit"] = None if peak_limit is None else np.broadcast_to(np.asarray(peak_limit, float), (T,)).copy()
    return P


def row_prox(v, lb, ub, rows, rho1, equality, mu0=None, iters=50):
    """z = argmin rho1/2 |z - v|^2 s.t. box, w*sum_win z <= (==) e.  Vectorised bisection."""
    z = np.clip(v, lb, ub)
    S = len(rows)
    if S == 0:
        return z, np.zeros(0)
    T = v.shape[1]
    ii = np.array([r[0] for r in rows]); aa = np.array([r[1] for r in rows]); bb = np.array([r[2] for r in rows])
    ww = np.array([r[3] for r in rows]); ee = np.array([r[4] for r in rows])
    tt = np.arange(T)[None, :]
    mask = (tt >= aa[:, None]) & (tt < bb[:, None])
    V = v[ii]; L = np.where(mask, lb[ii], 0.0); Ub = np.where(mask, ub[ii], 0.0)
    V = np.where(mask, V, 0.0)
    def E(mu):
        return ww * np.clip(V - (ww / rho1 * mu)[:, None], L, Ub).sum(axis=1)
    sc = rho1 / ww * (np.abs(V).max(axis=1) + np.abs(Ub).max(axis=1) + 1)
    lo = -sc if equality else np.zeros(S)
    hi = sc.copy()
    need = np.ones(S, bool) if equality else (E(np.zeros(S)) > ee)
    for _ in range(iters):
        mid = 0.5 * (lo + hi)
        g = E(mid) > ee
        lo = np.where(g, mid, lo); hi = np.where(g, hi, mid)
    mu = np.where(need, 0.5 * (lo + hi), 0.0)
    Z = np.clip(V - (ww / rho1 * mu)[:, None], L, Ub)
    for r in range(S):
        z[ii[r], aa[r]:bb[r]] = Z[r, aa[r]:bb[r]]
    return z, mu


def agg_prox(v, rho, Gamma, ebar, peaks):
    """argmin Gamma*sum (z+ebar)^2 + sum_c w_c max(max z, p0_c) + rho/2 |z-v|^2"""
    a = (rho * v - 2 * Gamma * ebar) / (rho + 2 * Gamma)
    if not peaks:
        return a
    cur = rho + 2 * Gamma
    def dphi(p):  # derivative of objective wrt level p (right derivative)
        return sum(w for w, p0 in peaks if p >= p0) - cur * np.maximum(a - p, 0).sum()
    lo, hi = a.min() - 1.0, a.max()
    pmin = min(p0 for _, p0 in peaks)
    # if at p = large derivative positive ... find root of dphi
    if dphi(hi) <= 0:
        return a
    lo = min(lo, pmin - 1)
    for _ in range(80):
        mid = 0.5 * (lo + hi)
        if dphi(mid) > 0:
            hi = mid
        else:
            lo = mid
    return np.minimum(a, hi)


def admm(P, rho=0.1, kappa=1.0, alpha=1.6, max_iter=5000, eps=1e-4, adapt=True, verbose=False, warm=None, check=10):
    N, T = P["N"], P["T"]
    lb, ub, c, qd, k = P["lb"], P["ub"], P["c"], P["qd"], P["k"]
    M = len(P["soc"]); ML = len(P["lin_rows"])
    has_pl = P["peak_limit"] is not None
    has_u = P["Gamma"] > 0 or len(P["peaks"]) > 0
    rowsK = []
    scales = []
    for j in range(M):
        s = np.sqrt((P["soc"][j] ** 2).sum() / 2) or 1.0
        rowsK += [P["soc"][j, 0] / s, P["soc"][j, 1] / s]; scales += [s, s]
    for j in range(ML):
        s = np.linalg.norm(P["lin_rows"][j]) or 1.0
        rowsK.append(P["lin_rows"][j] / s); scales.append(s)
    if has_pl:
        s = np.sqrt(N); rowsK.append(np.ones(N) / s); scales.append(s)
    if has_u:
        s = np.linalg.norm(k); rowsK.append(k / s); scales.append(s)
    K = np.array(rowsK).reshape(-1, N)
    scales = np.array(scales)
    R = len(K)
    KKt = K @ K.T
    lam, U = np.linalg.eigh(KKt) if R else (np.zeros(0), np.zeros((0, 0)))
    # cost scaling
    cs = 1.0 / max(np.abs(c).max(), 1e-12) if np.abs(c).max() > 0 else 1.0
    c = c * cs; qd = qd * cs; Gamma = P["Gamma"] * cs; peaks = [(w * cs, p0) for w, p0 in P["peaks"]]

    def proj_c(v, rho):
        z = v.copy()
        r = 0
        for j in range(M):
            lim = P["limits"][j] / scales[r]
            nrm = np.hypot(v[r], v[r + 1])
            f = np.minimum(1.0, lim / np.maximum(nrm, 1e-300))
            z[r] = v[r] * f; z[r + 1] = v[r + 1] * f
            r += 2
        for j in range(ML):
            z[r] = np.minimum(v[r], P["limits"][M + j if False else j] / scales[r]); r += 1
        if has_pl:
            z[r] = np.minimum(v[r], P["peak_limit"] / scales[r]); r += 1
        if has_u:
            su = scales[r]
            # variable is u/su; g(u) in terms of u = su*z
            zz = agg_prox(v[r] * su, rho / su**2, Gamma, P["ebar"], peaks)
            z[r] = zz / su; r += 1
        return z

    if warm is None:
        z1 = np.clip(np.zeros((N, T)), lb, ub); y1 = np.zeros((N, T)); zc = np.zeros((R, T)); yc = np.zeros((R, T))
    else:
        z1, y1, zc, yc = [w.copy() for w in warm]
    hist = []
    for it in range(1, max_iter + 1):
        rho1 = kappa * rho
        d = 2 * qd + rho1
        rhs = -c + rho1 * z1 - y1 + K.T @ (rho * zc - yc)
        w = K @ rhs
        if R:
            v = U @ ((U.T @ w) / (d / rho + lam)[:, None])
            x = (rhs - K.T @ v) / d
        else:
            x = rhs / d
        Kx = K @ x
        xa = alpha * x + (1 - alpha) * z1
        Ka = alpha * Kx + (1 - alpha) * zc
        v1 = xa + y1 / rho1
        z1n, mus = row_prox(v1, lb, ub, P["rows"], rho1, P["equality"])
        y1 = rho1 * (v1 - z1n)
        vc = Ka + yc / rho
        zcn = proj_c(vc, rho)
        yc = rho * (vc - zcn)
        dz1, dzc = z1n - z1, zcn - zc
        z1, zc = z1n, zcn
        if it % check == 0 or it == max_iter:
            rp = max(np.abs(x - z1).max(), np.abs(Kx - zc).max() if R else 0)
            rd = np.abs(rho1 * dz1 + rho * (K.T @ dzc)).max()
            pn = max(np.abs(x).max(), np.abs(z1).max(), 1e-9)
            dn = max(np.abs(c).max(), np.abs(y1).max(), np.abs(K.T @ yc).max() if R else 0, 1e-9)
            # violation of z1 wrt coupling constraints (true units, relative)
            Kz = K @ z1
            viol = 0.0
            r = 0
            for j in range(M):
                viol = max(viol, (np.hypot(Kz[r], Kz[r + 1]) * scales[r] / P["limits"][j] - 1).max()); r += 2
            for j in range(ML):
                viol = max(viol, (Kz[r] * scales[r] / P["limits"][j] - 1).max()); r += 1
            if has_pl:
                viol = max(viol, ((Kz[r] * scales[r] - P["peak_limit"]) / P["peak_limit"]).max()); r += 1
            hist.append((it, rp / pn, rd / dn, viol, rho))
            if verbose:
                print(it, f"rp {rp/pn:.2e} rd {rd/dn:.2e} viol {viol:.2e} rho {rho:.3g}")
            if rp / pn < eps and rd / dn < eps and viol < 1e-5:
                break
            if adapt and it % (check * 5) == 0:
                ratio = np.sqrt((rp / pn) / max(rd / dn, 1e-12))
                if ratio > 3 or ratio < 1 / 3:
                    rho_new = np.clip(rho * ratio, 1e-5, 1e5)
                    rho = rho_new
    return z1, x, it, hist, (z1, y1, zc, yc)


if __name__ == "__main__":
    from adacharge_b200.interface import TestingInterface
    from adacharge_b200.generators import *
    which = sys.argv[1] if len(sys.argv) > 1 else "c2"
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    if which == "c2":
        d = config_c2(seed=seed)
        obj = [("tou_energy_cost", 1, {}), ("total_energy", 1000, {}), ("demand_charge", 1, {})]
    else:
        d = config_c1(seed=seed)
        obj = [("quick_charge", 1, {}), ("equal_share", 1e-3, {})]
    iface = TestingInterface(d)
    S = iface.active_sessions(); I = iface.infrastructure_info()
    pp = iface.get_prev_peak()
    P = pack(obj, S, I, iface, prev_peak=pp)
    t = time.time()
    Ro, info = mpc.solve_mpc(obj, S, I, iface, prev_peak=pp, return_info=True)
    print("oracle", time.time() - t, info["iters"])
    fo = mpc.evaluate_objective(Ro, obj, I, iface)
    for rho in [0.01, 0.1, 1.0]:
        t = time.time()
        z, x, it, hist, _ = admm(P, rho=rho, verbose=False)
        f = mpc.evaluate_objective(z, obj, I, iface)
        print(f"rho0 {rho}: iters {it} obj {f:.6f} oracle {fo:.6f} rel {abs(f-fo)/abs(fo):.2e} viol", mpc.violations(z, S, I, iface), "maxdiff", np.abs(z - Ro).max(), f"{time.time()-t:.1f}s", hist[-1])
