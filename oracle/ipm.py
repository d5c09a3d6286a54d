"""Dense primal-dual interior-point solver for ``StageProblem`` batches.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).

Stands in for ``ca.nlpsol('ipopt')`` at the reference's call sites
(mpc_3d.py:82,124; np_mpc_adaptive_with_linear_regressor.py:158,214;
rlmpc2.py:491,508): a primal-dual log-barrier Newton method with the
fraction-to-boundary rule and a monotone barrier schedule (IPOPT's published
algorithm, Waechter & Biegler 2006; defaults mu_init=0.1, kappa_mu=0.2,
theta_mu=1.5, kappa_eps=10, tau_min=0.99, bound_push=0.01, tol=1e-8 from the
IPOPT documentation -- not from the reference tree).  It differs from the
CUDA solver in everything except the mathematics of the Newton step:

* the KKT system is assembled densely and solved with LAPACK (no Riccati
  recursion, no structure exploitation, no axis decoupling);
* first derivatives come from complex-step differentiation of the literal
  model restatements in ``oracle.models`` (no analytic Jacobians);
* everything is float64 numpy.

Because any solver converged to the same KKT tolerance finds the same point of
these small, locally convex NLPs, agreement between the two is the parity
statement; ``oracle.crosscheck`` provides a third, unrelated solver (SLSQP).
"""
from dataclasses import dataclass

import numpy as np

from .problems import StageProblem

STATUS_CONVERGED = 0
STATUS_MAXITER = 1
STATUS_INFEASIBLE = 2
STATUS_NUMERIC = 3
STATUS_ACCEPTABLE = 4   # IPOPT's acceptable-level termination (rlmpc2.py:486-488), off unless acc_iter > 0


@dataclass
class Options:
    tol: float = 1e-8
    max_iter: int = 200
    acc_tol: float = 0.0
    acc_iter: int = 0
    # multiplier scale of the cold start, z = mu0 / slack; None = the barrier strategy's default: 0.1 under the monotone
    # schedule (IPOPT's mu_init), 0.01 with predictor-corrector steps, which choose mu themselves (csrc: Solver::start_mu)
    mu0: float = None
    kappa_mu: float = 0.2
    theta_mu: float = 1.5
    kappa_eps: float = 10.0
    # experiments (tools/ipm_variants.py; the CUDA solver runs the monotone method): second-order corrector on the
    # complementarity products, and Mehrotra's adaptive barrier parameter from an affine-scaling predictor
    corrector: int = 0
    mehrotra: int = 0
    corr_passes: int = 1
    corr_mu: float = 1e300
    tau_min: float = 0.99
    bound_push: float = 1e-2
    eta: float = 1e-4
    smax: float = 100.0
    max_backtrack: int = 12
    s_phi: float = 2.3
    s_theta: float = 1.1
    delta_sw: float = 1.0
    gamma_theta: float = 1e-5
    gamma_phi: float = 1e-8
    theta_small: float = 1e-4
    curvature: bool = True


def jacobians(prob: StageProblem, X, U):
    """A_k = dF/dx, B_k = dF/du by the complex-step method (exact to rounding). -> F, A[B,N,n,n], Bm[B,N,n,m]."""
    n, m, N = prob.n, prob.m, prob.N
    h = 1e-30
    y = np.concatenate([X[:, :N], U], axis=-1).astype(np.complex128)          # [B,N,n+m]
    yy = y[:, :, None, :] + 1j * h * np.eye(n + m)[None, None]                # [B,N,n+m,n+m]
    Fp = prob.step(yy[..., :n], yy[..., n:])                                   # [B,N,n+m,n]
    Jt = np.imag(Fp) / h
    J = np.swapaxes(Jt, -1, -2)                                                # [B,N,n,n+m]
    F = prob.step(X[:, :N], U)
    return F, J[..., :n], J[..., n:]


def _row_mats(prob):
    """Dense C [nr, n+m] of the sparse rows."""
    C = np.zeros((len(prob.rows), prob.n + prob.m))
    for r, row in enumerate(prob.rows):
        C[r, row.ia] += row.sa
        if row.ib >= 0:
            C[r, row.ib] += row.sb
    return C


def cost_grad_hess(prob: StageProblem, X, U):
    """Gradient [B,N,n+m], terminal gradient [B,n], constant Hessians H[B,n+m,n+m], HT[B,n,n]."""
    n, m, N, B = prob.n, prob.m, prob.N, prob.B
    y = np.concatenate([X[:, :N], U], axis=-1)
    gy = 2.0 * prob.wy[:, None, :] * (y - prob.ry)
    H = np.zeros((B, n + m, n + m))
    idx = np.arange(n + m)
    H[:, idx, idx] = 2.0 * prob.wy
    if prob.naug:
        e = U - X[:, :N, n - m:]
        gy[:, :, n:] += 2.0 * prob.wd[:, None, :] * e
        gy[:, :, n - m:n] -= 2.0 * prob.wd[:, None, :] * e
        for j in range(m):
            a, b = n + j, n - m + j
            H[:, a, a] += 2.0 * prob.wd[:, j]
            H[:, b, b] += 2.0 * prob.wd[:, j]
            H[:, a, b] -= 2.0 * prob.wd[:, j]
            H[:, b, a] -= 2.0 * prob.wd[:, j]
    gT = 2.0 * prob.wT * (X[:, N] - prob.rT)
    HT = np.zeros((B, n, n))
    ii = np.arange(n)
    HT[:, ii, ii] = 2.0 * prob.wT
    return gy, gT, H, HT


def solve(prob: StageProblem, X0=None, U0=None, opts: Options = None, trace=None, warm=None):
    """Solve every instance of ``prob``. Returns dict(X, U, J, status, iters, lam, zl, zu, s, kkt).

    ``warm``: dict(lam, s, zl, zu, valid[B]) of a previous solve -- the dual warm start of dart_set_dual_state: where
    ``valid``, slacks are taken from it (pushed 1e-6 of the row range into the interior), bound multipliers kept above
    mu/(1e10 slack), equality multipliers as they are; elsewhere the default start, with mu not below 1e-4."""
    o = opts or Options()
    n, m, N, B = prob.n, prob.m, prob.N, prob.B
    nr = len(prob.rows)
    C = _row_mats(prob)                                   # [nr, n+m]
    msk = prob.row_mask()[None]                           # [1,N,nr]
    lo = np.stack([r.lo for r in prob.rows], axis=1)[:, None, :]   # [B,1,nr]
    hi = np.stack([r.hi for r in prob.rows], axis=1)[:, None, :]

    # ---- starting point (reference: tile(state) / zeros, mpc_3d.py:123; previous w0 for RMPC/LMPC)
    X = np.repeat(prob.x0[:, None, :], N + 1, axis=1) if X0 is None else np.array(X0, dtype=np.float64)
    U = np.zeros((B, N, m)) if U0 is None else np.array(U0, dtype=np.float64)
    X[:, 0] = prob.x0
    if prob.naug:
        X[:, 1:, n - m:] = U
    status = np.full(B, STATUS_MAXITER, dtype=np.int32)
    nacc = np.zeros(B, dtype=np.int64)
    infeasible0 = np.zeros(B, dtype=bool)
    for r, row in enumerate(prob.rows):       # rows skipped at k=0 must hold for the given x0
        if row.skip0:
            t0 = row.sa * prob.x0[:, row.ia]
            infeasible0 |= (t0 < row.lo) | (t0 > row.hi)

    # slacks pushed into the interior (IPOPT bound_push / bound_frac), u-box rows move u itself
    push = np.minimum(o.bound_push * np.maximum(1.0, np.maximum(np.abs(lo), np.abs(hi))), o.bound_push * (hi - lo))
    t = prob.row_values(X, U)
    s = np.clip(t, lo + push, hi - push)
    for r, row in enumerate(prob.rows):
        if row.ib < 0 and row.ia >= n:        # pure control bound: keep c.y == s exactly
            U[:, :, row.ia - n] = s[:, :, r] / row.sa
    if prob.naug:
        X[:, 1:, n - m:] = U
    mu = np.full(B, o.mu0 if o.mu0 is not None else (0.01 if o.mehrotra else 0.1))
    zl = (mu[:, None, None] / (s - lo)) * msk
    zu = (mu[:, None, None] / (hi - s)) * msk
    lam = np.zeros((B, N, n))                 # lam[:,k] multiplies F(x_k,u_k) - x_{k+1}
    if warm is not None:
        v = np.asarray(warm["valid"], dtype=bool)
        mu = np.where(v, mu, np.maximum(mu, 1e-4))
        zl = np.where(v[:, None, None], zl, (mu[:, None, None] / (s - lo)) * msk)
        zu = np.where(v[:, None, None], zu, (mu[:, None, None] / (hi - s)) * msk)
        p2 = 1e-6 * (hi - lo)
        sw = np.clip(warm["s"], lo + p2, hi - p2)
        s = np.where(v[:, None, None] & (msk > 0), sw, s)
        for r, row in enumerate(prob.rows):
            if row.ib < 0 and row.ia >= n:
                U[:, :, row.ia - n] = np.where(v[:, None], s[:, :, r] / row.sa, U[:, :, row.ia - n])
        if prob.naug:
            X[:, 1:, n - m:] = U
        zl = np.where(v[:, None, None], np.maximum(warm["zl"], mu[:, None, None] / (1e10 * (s - lo))) * msk, zl)
        zu = np.where(v[:, None, None], np.maximum(warm["zu"], mu[:, None, None] / (1e10 * (hi - s))) * msk, zu)
        lam = np.where(v[:, None, None], warm["lam"], lam)
    iters = np.zeros(B, dtype=np.int32)
    tiny = np.zeros(B, dtype=np.int32)
    ls_evals = np.zeros(B, dtype=np.int64)     # trial points evaluated by the line searches (cost statistics for the tools)
    done = np.zeros(B, dtype=bool)
    nv = N * (m + n)
    ne = N * n
    kkt_final = np.zeros(B)
    nrows_act = float(msk.sum())

    def iu(k):
        return slice(k * (m + n), k * (m + n) + m)

    def ix(k):        # k >= 1
        return slice((k - 1) * (m + n) + m, k * (m + n))

    for it in range(o.max_iter + 1):
        F, A, Bm = jacobians(prob, X, U)
        d = F - X[:, 1:]
        gy, gT, Hc, HT = cost_grad_hess(prob, X, U)
        t = prob.row_values(X, U)
        rc = (t - s) * msk
        sl, su = s - lo, hi - s
        nu = (zu - zl) * msk
        # ---- KKT residuals
        Cnu = nu @ C                                        # [B,N,n+m]
        gL = gy + Cnu
        gL[:, :, :n] += np.einsum('bkij,bki->bkj', A, lam)
        gL[:, :, n:] += np.einsum('bkij,bki->bkj', Bm, lam)
        gL[:, 1:, :n] -= lam[:, :-1]
        gL[:, 0, :n] = 0.0                                  # x_0 is not a variable
        gLT = gT - lam[:, -1]
        dual_inf = np.maximum(np.abs(gL).max(axis=(1, 2)), np.abs(gLT).max(axis=1))
        prim_inf = np.maximum(np.abs(d).max(axis=(1, 2)), np.abs(rc).max(axis=(1, 2)))
        zsum = (np.abs(zl) + np.abs(zu)).sum(axis=(1, 2))
        s_d = np.maximum(o.smax, (np.abs(lam).sum(axis=(1, 2)) + zsum) / (ne + 2 * nrows_act)) / o.smax
        s_c = np.maximum(o.smax, zsum / (2 * nrows_act)) / o.smax

        def compl(mu_):
            a = np.abs(zl * sl - mu_[:, None, None]) * msk
            b = np.abs(zu * su - mu_[:, None, None]) * msk
            return np.maximum(a.max(axis=(1, 2)), b.max(axis=(1, 2)))

        E0 = np.maximum(np.maximum(dual_inf / s_d, prim_inf), compl(np.zeros(B)) / s_c)
        newly = (~done) & (E0 <= o.tol)
        status[newly] = STATUS_CONVERGED
        kkt_final = np.where(done, kkt_final, E0)
        done |= newly
        bad = (~done) & ~np.isfinite(E0)
        status[bad] = STATUS_NUMERIC
        done |= bad
        if o.acc_iter > 0:      # consecutive acceptable iterates
            nacc = np.where((~done) & (E0 <= o.acc_tol), nacc + 1, 0)
            acc = (~done) & (nacc >= o.acc_iter)
            status[acc] = STATUS_ACCEPTABLE
            done |= acc
        if trace is not None:
            trace.append(dict(it=it, E0=E0.copy(), mu=mu.copy(), dual=dual_inf.copy(), prim=prim_inf.copy(),
                              J=prob.objective(X, U), done=done.copy()))
        if done.all() or it == o.max_iter:
            break
        iters[~done] += 1
        # ---- monotone barrier update (Waechter & Biegler eq. 7), possibly several reductions at once
        mu_min = o.tol / 10.0
        for _ in range(0 if o.mehrotra else 8):
            Emu = np.maximum(np.maximum(dual_inf / s_d, prim_inf), compl(mu) / s_c)
            red = (~done) & (Emu <= o.kappa_eps * mu) & (mu > mu_min)
            if not red.any():
                break
            mu = np.where(red, np.maximum(mu_min, np.minimum(o.kappa_mu * mu, mu ** o.theta_mu)), mu)
        if o.mehrotra:
            cnt = 2.0 * nrows_act
            mu_c = ((zl * sl + zu * su) * msk).sum(axis=(1, 2)) / cnt      # mean complementarity product
            mu = np.zeros(B)                                               # predictor: affine-scaling direction
        mu3 = mu[:, None, None]
        # ---- condensed Newton system
        Sig = (zl / sl + zu / su) * msk
        nuhat = (mu3 / su - mu3 / sl + Sig * rc) * msk
        Hs = Hc[:, None] + np.einsum('bkr,ri,rj->bkij', Sig, C, C)         # [B,N,n+m,n+m]
        if o.curvature:
            # Lagrangian curvature of the tilt input: the models are control-affine in g*sin(u_i), so
            # d2(lam.F)/du_i^2 = -tan(u_i) * (B^T lam)_i (exact for PMPC, O(Ts^2) off for RMPC/LMPC).
            Bl = np.einsum('bkij,bki->bkj', Bm, lam)
            for j in range(m):
                Hs[:, :, n + j, n + j] += -np.tan(U[:, :, j]) * Bl[:, :, j]
        gs = gy + nuhat @ C
        delta_w = np.zeros(B)
        for attempt in range(12):
            K = np.zeros((B, nv + ne, nv + ne))
            rhs = np.zeros((B, nv + ne))
            reg = delta_w[:, None, None] * np.eye(n + m)[None]
            for k in range(N):
                Hk = Hs[:, k] + reg
                K[:, iu(k), iu(k)] += Hk[:, n:, n:]
                rhs[:, iu(k)] -= gs[:, k, n:]
                if k >= 1:
                    K[:, ix(k), ix(k)] += Hk[:, :n, :n]
                    K[:, ix(k), iu(k)] += Hk[:, :n, n:]
                    K[:, iu(k), ix(k)] += Hk[:, n:, :n]
                    rhs[:, ix(k)] -= gs[:, k, :n]
                er = slice(nv + k * n, nv + (k + 1) * n)
                K[:, er, iu(k)] += Bm[:, k]
                K[:, iu(k), er] += np.swapaxes(Bm[:, k], 1, 2)
                if k >= 1:
                    K[:, er, ix(k)] += A[:, k]
                    K[:, ix(k), er] += np.swapaxes(A[:, k], 1, 2)
                K[:, er, ix(k + 1)] -= np.eye(n)[None]
                K[:, ix(k + 1), er] -= np.eye(n)[None]
                rhs[:, er] = -d[:, k]
            K[:, ix(N), ix(N)] += HT + delta_w[:, None, None] * np.eye(n)[None]
            rhs[:, ix(N)] -= gT
            sol = np.linalg.solve(K, rhs[..., None])[..., 0]
            dv = sol[:, :nv].reshape(B, N, m + n)
            dU = dv[:, :, :m]
            dX = np.zeros((B, N + 1, n))
            dX[:, 1:] = dv[:, :, m:]
            lam_new = sol[:, nv:].reshape(B, N, n)
            # curvature test on the step: d^T H d must be positive where constraints are linearly satisfied
            dy = np.concatenate([dX[:, :N], dU], axis=-1)
            quad = np.einsum('bki,bkij,bkj->b', dy, Hs + reg[:, None], dy) + \
                np.einsum('bi,bij,bj->b', dX[:, N], HT + delta_w[:, None, None] * np.eye(n)[None], dX[:, N])
            need = (~done) & (quad <= 0.0) & (np.abs(dy).max(axis=(1, 2)) > 0)
            if not need.any():
                break
            delta_w = np.where(need, np.where(delta_w == 0.0, 1e-4, delta_w * 8.0), delta_w)
        ds = (dy @ C.T + rc) * msk
        dzl = (mu3 / sl - zl - (zl / sl) * ds) * msk
        dzu = (mu3 / su - zu + (zu / su) * ds) * msk
        if o.mehrotra:
            # barrier parameter from the predictor: mu = sigma * mean(z s), sigma = (mean after the longest affine step / mean)^3
            with np.errstate(divide='ignore', invalid='ignore'):
                a1 = np.where(ds < 0, -sl / ds, np.inf)
                a2 = np.where(ds > 0, su / ds, np.inf)
                apa = np.minimum(1.0, np.minimum(np.where(msk > 0, a1, np.inf).min(axis=(1, 2)),
                                                 np.where(msk > 0, a2, np.inf).min(axis=(1, 2))))
                b1 = np.where(dzl < 0, -zl / dzl, np.inf)
                b2 = np.where(dzu < 0, -zu / dzu, np.inf)
                ada = np.minimum(1.0, np.minimum(np.where(msk > 0, b1, np.inf).min(axis=(1, 2)),
                                                 np.where(msk > 0, b2, np.inf).min(axis=(1, 2))))
            pa, da = apa[:, None, None], ada[:, None, None]
            mu_aff = (((sl + pa * ds) * (zl + da * dzl) + (su - pa * ds) * (zu + da * dzu)) * msk).sum(axis=(1, 2)) / cnt
            mu = np.maximum(np.clip((mu_aff / mu_c) ** 3, 1e-8, 1.0) * mu_c, o.tol / 10.0)
            mu3 = mu[:, None, None]
        if o.corrector or o.mehrotra:
            # second solve with the same matrix: (s + ds)(z + dz) = mu keeps the product ds dz of the first direction
            for _pass in range(max(1, o.corr_passes)):      # > 1: the products of the corrected direction, solved again (experiment)
                if _pass > 0:
                    keep = (dX.copy(), dU.copy(), lam_new.copy(), ds.copy(), dzl.copy(), dzu.copy())
                ml = mu3 - ds * dzl
                mu_ = mu3 + ds * dzu
                gs2 = gy + ((mu_ / su - ml / sl + Sig * rc) * msk) @ C
                for k in range(N):
                    rhs[:, iu(k)] = -gs2[:, k, n:]
                    if k >= 1:
                        rhs[:, ix(k)] = -gs2[:, k, :n]
                sol = np.linalg.solve(K, rhs[..., None])[..., 0]
                dv = sol[:, :nv].reshape(B, N, m + n)
                dU = dv[:, :, :m]
                dX = np.zeros((B, N + 1, n))
                dX[:, 1:] = dv[:, :, m:]
                lam_new = sol[:, nv:].reshape(B, N, n)
                dy = np.concatenate([dX[:, :N], dU], axis=-1)
                ds = (dy @ C.T + rc) * msk
                dzl = (ml / sl - zl - (zl / sl) * ds) * msk
                dzu = (mu_ / su - zu + (zu / su) * ds) * msk
                if _pass > 0:       # further passes only where the barrier parameter is already small (the linear tail)
                    use = mu <= o.corr_mu
                    u3 = use[:, None, None]
                    dX = np.where(u3, dX, keep[0]); dU = np.where(u3, dU, keep[1]); lam_new = np.where(u3, lam_new, keep[2])
                    ds = np.where(u3, ds, keep[3]); dzl = np.where(u3, dzl, keep[4]); dzu = np.where(u3, dzu, keep[5])
        # ---- fraction to the boundary
        tau = np.maximum(o.tau_min, 1.0 - mu)[:, None, None]
        with np.errstate(divide='ignore', invalid='ignore'):
            a1 = np.where(ds < 0, -tau * sl / ds, np.inf)
            a2 = np.where(ds > 0, tau * su / ds, np.inf)
            ap = np.minimum(1.0, np.minimum(np.where(msk > 0, a1, np.inf).min(axis=(1, 2)),
                                            np.where(msk > 0, a2, np.inf).min(axis=(1, 2))))
            b1 = np.where(dzl < 0, -tau * zl / dzl, np.inf)
            b2 = np.where(dzu < 0, -tau * zu / dzu, np.inf)
            ad = np.minimum(1.0, np.minimum(np.where(msk > 0, b1, np.inf).min(axis=(1, 2)),
                                            np.where(msk > 0, b2, np.inf).min(axis=(1, 2))))
        # ---- filter-type acceptance (Waechter & Biegler sec. 2.3, without filter history): a trial point
        # is accepted if it gives Armijo decrease of the barrier objective when the switching condition
        # holds, else if it sufficiently reduces either the constraint violation or the barrier objective.
        def barrier_obj(X_, U_, s_):
            with np.errstate(invalid='ignore', divide='ignore'):
                bar = -(mu3 * (np.log(s_ - lo) + np.log(hi - s_)) * msk).sum(axis=(1, 2))
            F_ = prob.step(X_[:, :N], U_)
            viol = np.abs(F_ - X_[:, 1:]).sum(axis=(1, 2)) + np.abs((prob.row_values(X_, U_) - s_) * msk).sum(axis=(1, 2))
            return prob.objective(X_, U_) + bar, viol

        phi0, th0 = barrier_obj(X, U, s)
        Dphi = (gy * dy).sum(axis=(1, 2)) + (gT * dX[:, N]).sum(axis=1) \
            - (mu3 * ds * (1.0 / sl - 1.0 / su) * msk).sum(axis=(1, 2))
        th_max = 1e4 * np.maximum(1.0, th0)
        alpha = ap.copy()
        accepted = done.copy()
        for _ in range(o.max_backtrack):
            ls_evals += ~accepted
            a3 = alpha[:, None, None]
            Xt, Ut, st = X + a3 * dX, U + a3 * dU, s + a3 * ds
            phit, tht = barrier_obj(Xt, Ut, st)
            switching = (Dphi < 0) & (alpha * np.abs(Dphi) ** o.s_phi > o.delta_sw * th0 ** o.s_theta)
            armijo = phit <= phi0 + o.eta * alpha * Dphi + 10 * np.finfo(float).eps * np.abs(phi0)
            suff = (tht <= (1 - o.gamma_theta) * th0) | (phit <= phi0 - o.gamma_phi * th0)
            ok = np.where(switching & (th0 <= o.theta_small), armijo, suff) & (tht <= th_max) & np.isfinite(phit)
            accepted |= ok
            if accepted.all():
                break
            alpha = np.where(accepted, alpha, alpha * 0.5)
        # the step vanished three times in a row (no restoration phase): stop, flag infeasible if violation remains
        tiny = np.where((~done) & (alpha <= 1e-6), tiny + 1, 0)
        stall = (~done) & (tiny >= 3)
        upd = (~done)
        a3 = np.where(upd, alpha, 0.0)[:, None, None]
        X = X + a3 * dX
        U = U + a3 * dU
        s = s + a3 * ds
        lam = lam + a3 * (lam_new - lam)
        ad3 = np.where(upd, ad, 0.0)[:, None, None]
        zl = zl + ad3 * dzl
        zu = zu + ad3 * dzu
        # IPOPT eq. (16): keep z within [mu/(k s), k mu/s], k = 1e10
        ks = 1e10
        sl, su = s - lo, hi - s
        zl = np.where(msk > 0, np.clip(zl, mu3 / (ks * sl), ks * mu3 / sl), 0.0)
        zu = np.where(msk > 0, np.clip(zu, mu3 / (ks * su), ks * mu3 / su), 0.0)
        if stall.any():
            Fn = prob.step(X[:, :N], U)
            pin = np.maximum(np.abs(Fn - X[:, 1:]).max(axis=(1, 2)), np.abs((prob.row_values(X, U) - s) * msk).max(axis=(1, 2)))
            status[stall] = np.where(pin[stall] > 1e-4, STATUS_INFEASIBLE, STATUS_MAXITER)
            done |= stall

    status = np.where(infeasible0 & (status != STATUS_NUMERIC), STATUS_INFEASIBLE, status)
    return dict(X=X, U=U, J=prob.objective(X, U), status=status, iters=iters, lam=lam, zl=zl, zu=zu, s=s,
                kkt=kkt_final, mu=mu, ls_evals=ls_evals)
