"""Dense primal-dual interior-point solver for a captured NLP tape.  TEST INFRASTRUCTURE ONLY.

    min f(x, p)   s.t.  lbg <= g(x, p) <= ubg,   lbx <= x <= ubx

``tape`` has inputs x, p and outputs f, g (oracle.refshim.casadi.NlpSolver).  All derivatives are exact to rounding:
gradient and constraint Jacobian from the tape's reverse sweep / complex step, the Hessian of the Lagrangian by a complex
step over the reverse sweep.  Method: IPOPT's published algorithm in its plainest form (Waechter & Biegler 2006) --
slack variables for inequality rows, log barrier on all bounds, monotone barrier decrease, fraction-to-the-boundary
rule, inertia correction by a diagonal shift, a filter line search -- solved to a much tighter
tolerance (1e-10 scaled KKT error) than any of the reference's IPOPT settings, because its output is used as a golden
KKT point, not as a timing stand-in.  It shares no code with ``oracle.ipm`` (which exploits the stage structure) or
with the CUDA solver.
"""
import numpy as np


def solve(tape, x0, p, lbx, ubx, lbg, ubg, tol=1e-10, max_iter=300, mu0=0.1, verbose=False):
    n = len(x0)
    m = len(lbg)
    eq = np.isfinite(lbg) & np.isfinite(ubg) & (lbg == ubg)
    iq = ~eq & (np.isfinite(lbg) | np.isfinite(ubg))
    ie, ii = np.nonzero(eq)[0], np.nonzero(iq)[0]
    ns = len(ii)
    nv = n + ns
    vL = np.concatenate([lbx, lbg[ii]])
    vU = np.concatenate([ubx, ubg[ii]])
    hasL, hasU = np.isfinite(vL), np.isfinite(vU)
    nc = len(ie) + ns

    def funcs(x):
        ev = tape.eval(x=x, p=p)
        return float(ev["f"][0]), ev["g"]

    def push_inside(v):
        v = v.copy()
        k1 = k2 = 1e-2
        pl = np.where(hasL, np.minimum(k1 * np.maximum(1.0, np.abs(np.where(hasL, vL, 0.0))), k2 * np.where(hasL & hasU, vU - vL, np.inf)), 0.0)
        pu = np.where(hasU, np.minimum(k1 * np.maximum(1.0, np.abs(np.where(hasU, vU, 0.0))), k2 * np.where(hasL & hasU, vU - vL, np.inf)), 0.0)
        v = np.where(hasL, np.maximum(v, vL + pl), v)
        v = np.where(hasU, np.minimum(v, vU - pu), v)
        return v

    x = np.asarray(x0, float).copy()
    f, g = funcs(x)
    v = push_inside(np.concatenate([x, g[ii]]))
    zL = np.where(hasL, 1.0, 0.0)
    zU = np.where(hasU, 1.0, 0.0)
    lam = np.zeros(nc)
    mu = mu0
    filt, theta_ref = [], None
    status, it = 1, 0
    delta_last = 0.0
    kkt = np.inf

    def cons(g, s):
        return np.concatenate([g[ie] - lbg[ie], g[ii] - s])

    def barrier(v):
        b = 0.0
        if hasL.any():
            b -= np.sum(np.log(v[hasL] - vL[hasL]))
        if hasU.any():
            b -= np.sum(np.log(vU[hasU] - v[hasU]))
        return b

    for it in range(max_iter + 1):
        x, s = v[:n], v[n:]
        f, g = funcs(x)
        c = cons(g, s)
        gf = tape.grad("f", "x", x=x, p=p)
        Jg = tape.jac("g", "x", x=x, p=p)                               # [m, n]
        A = np.zeros((nc, nv))
        A[:len(ie), :n] = Jg[ie]
        A[len(ie):, :n] = Jg[ii]
        A[len(ie):, n:] = -np.eye(ns)
        gv = np.concatenate([gf, np.zeros(ns)])
        dL = np.where(hasL, v - vL, 1.0)
        dU = np.where(hasU, vU - v, 1.0)
        r_dual = gv + A.T @ lam - zL + zU
        compL = np.where(hasL, dL * zL, 0.0)
        compU = np.where(hasU, dU * zU, 0.0)

        def err(mu_):
            sd = max(100.0, (np.abs(lam).sum() + zL.sum() + zU.sum()) / max(1, nc + hasL.sum() + hasU.sum())) / 100.0
            sc = max(100.0, (zL.sum() + zU.sum()) / max(1, hasL.sum() + hasU.sum())) / 100.0
            e3 = 0.0
            if hasL.any():
                e3 = max(e3, np.abs(compL[hasL] - mu_).max())
            if hasU.any():
                e3 = max(e3, np.abs(compU[hasU] - mu_).max())
            return max(np.abs(r_dual).max() / sd, np.abs(c).max() if nc else 0.0, e3 / sc)

        kkt = err(0.0)
        if verbose:
            print(f"it {it:3d} f={f:.10g} |c|={np.abs(c).max() if nc else 0:.2e} dual={np.abs(r_dual).max():.2e} mu={mu:.1e} kkt={kkt:.2e}")
        if kkt <= tol:
            status = 0
            break
        if it == max_iter:
            break
        while err(mu) <= 10.0 * mu and mu > tol / 10.0:
            mu = max(tol / 10.0, min(0.2 * mu, mu ** 1.5))
            filt = []

        lam_g_full = np.zeros(m)
        lam_g_full[ie] = lam[:len(ie)]
        lam_g_full[ii] = lam[len(ie):]
        W = np.zeros((nv, nv))
        W[:n, :n] = tape.hess_lagrangian({"f": np.ones(1), "g": lam_g_full}, "x", x=x, p=p)
        Sig = np.where(hasL, zL / dL, 0.0) + np.where(hasU, zU / dU, 0.0)
        rhs_v = -(gv + A.T @ lam - np.where(hasL, mu / dL, 0.0) + np.where(hasU, mu / dU, 0.0))
        rhs = np.concatenate([rhs_v, -c])
        delta = 0.0
        while True:
            K = np.zeros((nv + nc, nv + nc))
            K[:nv, :nv] = W + np.diag(Sig + delta)
            K[:nv, nv:] = A.T
            K[nv:, :nv] = A
            npos, nneg = _inertia(K)
            if npos == nv and nneg == nc:
                break
            delta = max(1e-4, delta_last / 3.0) if delta == 0.0 else delta * 8.0
            if delta > 1e20:
                return _result(tape, x, p, f, g, lam, zL, zU, ie, ii, n, m, it, kkt, 3)
        if delta > 0:
            delta_last = delta
        sol = np.linalg.solve(K, rhs)
        sol += np.linalg.solve(K, rhs - K @ sol)                        # one step of iterative refinement
        dv, dlam = sol[:nv], sol[nv:]
        dzL = np.where(hasL, mu / dL - zL - zL / dL * dv, 0.0)
        dzU = np.where(hasU, mu / dU - zU + zU / dU * dv, 0.0)
        tau = max(0.99, 1.0 - mu)

        def max_step(val, dval, mask):
            k = mask & (dval < 0)
            return min(1.0, float((-tau * val[k] / dval[k]).min())) if k.any() else 1.0

        a_pr = min(max_step(dL, dv, hasL), max_step(dU, -dv, hasU))
        a_du = min(max_step(zL, dzL, hasL), max_step(zU, dzU, hasU))

        # filter line search (Waechter & Biegler 2006, section 2.3) without a restoration phase
        th0 = np.abs(c).sum()
        phi0 = f + mu * barrier(v)
        dphi = gv @ dv - mu * np.sum(np.where(hasL, dv / dL, 0.0)) + mu * np.sum(np.where(hasU, dv / dU, 0.0))
        if theta_ref is None:
            theta_ref = max(1.0, th0)
        th_min, th_max = 1e-4 * theta_ref, 1e4 * theta_ref
        def acceptable(vt, a):
            ft, gt = funcs(vt[:n])
            if not (np.isfinite(ft) and np.all(np.isfinite(gt))):
                return False, np.inf
            tht = np.abs(cons(gt, vt[n:])).sum()
            pht = ft + mu * barrier(vt)
            if tht > th_max or any(tht >= tf and pht >= pf for tf, pf in filt):
                return False, tht
            if th0 <= th_min and dphi < 0 and a * (-dphi) ** 2.3 > th0 ** 1.1:            # switching condition: Armijo on phi
                return pht <= phi0 + 1e-4 * a * dphi + 1e-13 * max(1.0, abs(phi0)), tht
            if tht <= (1 - 1e-5) * th0 or pht <= phi0 - 1e-5 * th0 + 1e-13 * max(1.0, abs(phi0)):
                filt.append(((1 - 1e-5) * th0, phi0 - 1e-5 * th0))
                return True, tht
            return False, tht

        a = a_pr
        accepted = False
        for ls in range(40):
            vt = v + a * dv
            accepted, tht = acceptable(vt, a)
            if accepted:
                break
            if ls == 0 and tht >= th0 and nc:
                # second-order correction: re-solve with the constraint residual of the trial point folded in
                c_soc = a * c
                th_old = tht
                for _soc in range(4):
                    ft, gt = funcs(vt[:n])
                    c_soc = a * c_soc + cons(gt, vt[n:]) if _soc else a * c + cons(gt, vt[n:])
                    sol2 = np.linalg.solve(K, np.concatenate([rhs_v, -c_soc]))
                    dv2 = sol2[:nv]
                    a2 = min(max_step(dL, dv2, hasL), max_step(dU, -dv2, hasU))
                    vt = v + a2 * dv2
                    ok2, tht2 = acceptable(vt, a2)
                    if ok2:
                        accepted, a, dv, dlam = True, a2, dv2, sol2[nv:]
                        break
                    if not np.isfinite(tht2) or tht2 > 0.99 * th_old:
                        break
                    th_old = tht2
                    a = a2
                if accepted:
                    break
                a = a_pr
            a *= 0.5
        if not accepted:
            a = a_pr * 1e-3                                       # take a short step rather than stall; errors show up in kkt
            vt = v + a * dv
        v = vt
        lam = lam + a * dlam
        zL = np.where(hasL, zL + a_du * dzL, 0.0)
        zU = np.where(hasU, zU + a_du * dzU, 0.0)
        # keep the bound multipliers within IPOPT's kappa_Sigma band around mu / slack
        dL = np.where(hasL, v - vL, 1.0)
        dU = np.where(hasU, vU - v, 1.0)
        zL = np.where(hasL, np.clip(zL, mu / (1e10 * dL), 1e10 * mu / dL), 0.0)
        zU = np.where(hasU, np.clip(zU, mu / (1e10 * dU), 1e10 * mu / dU), 0.0)

    x = v[:n]
    f, g = funcs(x)
    return _result(tape, x, p, f, g, lam, zL, zU, ie, ii, n, m, it, kkt, status)


def _inertia(K):
    """(#positive, #negative) eigenvalues from a Bunch-Kaufman LDL^T factorisation (Sylvester's law), as IPOPT reads it
    off its linear solver."""
    from scipy.linalg import ldl
    _, D, _ = ldl(K)
    npos = nneg = 0
    i, nK = 0, K.shape[0]
    while i < nK:
        if i + 1 < nK and D[i + 1, i] != 0.0:
            w = np.linalg.eigvalsh(D[i:i + 2, i:i + 2])
            npos += int((w > 0).sum())
            nneg += int((w < 0).sum())
            i += 2
        else:
            npos += int(D[i, i] > 0)
            nneg += int(D[i, i] < 0)
            i += 1
    return npos, nneg


def _result(tape, x, p, f, g, lam, zL, zU, ie, ii, n, m, it, kkt, status):
    lam_g = np.zeros(m)
    lam_g[ie] = lam[:len(ie)]
    lam_g[ii] = lam[len(ie):]
    lam_x = zU[:n] - zL[:n]
    return {"x": x.copy(), "f": float(f), "g": np.asarray(g).copy(), "lam_g": lam_g, "lam_x": lam_x,
            "iters": int(it), "kkt": float(kkt), "status": int(status)}
