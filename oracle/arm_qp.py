"""CPU oracle for the low-level arm QP.  TEST INFRASTRUCTURE ONLY.

Restates ``PMPC/src/controller/arm.py:337-405`` (``ARMCONTROL.solver_worker``): every 2 ms, per arm, a 7-variable
problem in the joint accelerations ``qdd``

    min  Eimp' Wimp Eimp + Epos' Wpos Epos + qddd' Wsmooth qddd
         Eimp = J qdd + Jdot qd - Mx_inv F,   F = -D (J qd) + K twist + mu
         Epos = qdd - beta,                   beta = 2 sqrt(diag K_null) (-qd) + K_null (-q)
         qddd = (qdd - qdd_prev) / dt
    s.t. Qmin    <= q + qd dt + 0.5 qdd dt^2 <= Qmax
         Qdotmin <= qd + qdd dt              <= Qdotmax
         taumin  <= M qdd + h                <= taumax

which the reference hands to CasADi/IPOPT as a general NLP.  It is a strictly convex QP (Wpos = 0.1 I), so its minimiser
is unique: any correct solver must return the same ``qdd`` -- that is what pins parity here, together with the KKT
residual check in ``kkt_residual`` and the scipy cross-check in ``tests/test_oracle_arm_qp.py``.  CasADi itself is not
installable (see ``oracle/__init__.py``): PARITY UNPINNED against the reference's floating-point output, pinned against
the mathematical optimum.

``build_qp`` follows the reference line by line (numpy, one instance at a time, as the reference does);
``solve_qp`` is a dense primal-dual interior-point method (slack form, monotone barrier, fraction to the boundary) that
shares no code with the CUDA kernel.
"""
from dataclasses import dataclass

import numpy as np

NV, NR = 7, 21

STATUS_CONVERGED, STATUS_MAXITER, STATUS_INFEASIBLE, STATUS_NUMERIC = 0, 1, 2, 3


def default_params(dt=0.002):
    """arm.py:495-519 (R_params; L_params carry the same numbers)."""
    return dict(
        Wimp=np.diag([10.0, 10.0, 10.0, 1.0, 1.0, 1.0]), Wpos=np.eye(7) * 0.1, Wsmooth=np.eye(7) * 0.0,
        Qmin=np.array([-6.28319, -2.059, -6.28319, -0.19198, -6.28319, -1.69297, -6.28319]),
        Qmax=np.array([6.28319, 2.0944, 6.28319, 3.927, 6.28319, 3.14159, 6.28319]),
        Qdotmin=np.ones(7) * -20.0, Qdotmax=np.ones(7) * 20.0,
        taumin=np.array([-50.0, -50, -30, -30, -30, -20, -20]), taumax=np.array([50.0, 50, 30, 30, 30, 20, 20]),
        K=np.diag([1000.0, 1000.0, 1000.0, 50.0, 50.0, 50.0]) * 10, K_null=np.diag([1.0] * 7), dt=dt)


def safe_matrix_sqrt(matrix):
    """arm.py:363-366."""
    eigvals, eigvecs = np.linalg.eigh(matrix)
    return eigvecs @ np.diag(np.sqrt(np.abs(eigvals))) @ eigvecs.T


def build_qp_one(dyn, params):
    """One instance: arm.py:337-405.  Returns H, g, c0 (cost = 0.5 x'Hx + g'x + c0), C [21,7], lo, hi."""
    q, qd, qdd_prev = dyn["q"], dyn["qd"], dyn["qdd_prev"]
    jac, jacDot, M, h, Mx_inv = dyn["jac"], dyn["jacDot"], dyn["M"], dyn["h"], dyn["Mx_inv"]
    K, K_null, dt = params["K"], params["K_null"], params["dt"]
    dx = dyn["mocap_pos"] - dyn["ee_pos"]
    twist = np.zeros(6)
    twist[:3] = dx
    twist[3:] = dyn["rotvec"]
    Minv = np.linalg.pinv(M, rcond=1e-6)                                        # :346-349
    if abs(np.linalg.det(Mx_inv)) > 1e-8:                                        # :351-357
        Mx = np.linalg.inv(Mx_inv)
    else:
        Mx = np.linalg.pinv(Mx_inv, rcond=1e-3)
    mu_np = Mx @ (jac @ (Minv @ h) + jacDot @ qd)                                # :360
    D_np = safe_matrix_sqrt(Mx) @ np.sqrt(K) + np.sqrt(K) @ safe_matrix_sqrt(Mx)   # :368-370
    F = -D_np @ (jac @ qd) + K @ twist + mu_np                                   # :384
    e0 = jacDot @ qd - Mx_inv @ F                                                # Eimp = jac qdd + e0, :385
    beta = 2.0 * np.sqrt(np.diag(K_null)) * (-qd) + (K_null @ (-q))              # :387
    Wimp, Wpos, Wsm = params["Wimp"], params["Wpos"], params["Wsmooth"] / dt ** 2
    H = 2.0 * (jac.T @ Wimp @ jac + Wpos + Wsm)
    H = 0.5 * (H + H.T)
    g = 2.0 * (jac.T @ (Wimp @ e0) - Wpos @ beta - Wsm @ qdd_prev)
    c0 = e0 @ Wimp @ e0 + beta @ Wpos @ beta + qdd_prev @ Wsm @ qdd_prev
    C = np.vstack([0.5 * dt ** 2 * np.eye(7), dt * np.eye(7), M])               # :399-402
    off = np.concatenate([qd * dt + q, qd, h])
    lo = np.concatenate([params["Qmin"], params["Qdotmin"], params["taumin"]]) - off
    hi = np.concatenate([params["Qmax"], params["Qdotmax"], params["taumax"]]) - off
    return H, g, c0, C, lo, hi


def build_qp(dyn, params):
    """Batched: every entry of ``dyn`` has a leading axis B."""
    B = dyn["q"].shape[0]
    out = [build_qp_one({k: v[b] for k, v in dyn.items()}, params) for b in range(B)]
    return tuple(np.stack([o[i] for o in out]) for i in range(6))


@dataclass
class QPOptions:
    tol: float = 1e-8
    max_iter: int = 100
    mu0: float = 0.1
    kappa_mu: float = 0.2
    theta_mu: float = 1.5
    kappa_eps: float = 10.0
    tau_min: float = 0.99
    bound_push: float = 1e-2
    smax: float = 100.0


def solve_qp(H, g, C, lo, hi, x0=None, opts=None):
    """min 0.5 x'Hx + g'x  s.t. lo <= Cx <= hi, batched over the leading axis.  Slack form Cx - s = 0, lo <= s <= hi."""
    o = opts or QPOptions()
    B, nr, nv = C.shape
    x = np.zeros((B, nv)) if x0 is None else np.array(x0, dtype=float)
    t = np.einsum('brv,bv->br', C, x)
    push = np.minimum(o.bound_push * np.maximum(1.0, np.maximum(np.abs(lo), np.abs(hi))), o.bound_push * (hi - lo))
    s = np.minimum(np.maximum(t, lo + push), hi - push)
    mu = np.full(B, o.mu0)
    zl = mu[:, None] / (s - lo)
    zu = mu[:, None] / (hi - s)
    status = np.full(B, STATUS_MAXITER, dtype=np.int32)
    iters = np.zeros(B, dtype=np.int32)
    done = (hi - lo <= 0).any(axis=1)                       # empty box: no interior
    status[done] = STATUS_INFEASIBLE
    mu_min = o.tol / 10.0
    for it in range(o.max_iter + 1):
        sl, su = s - lo, hi - s
        rc = np.einsum('brv,bv->br', C, x) - s
        nu = zu - zl
        grad = np.einsum('bij,bj->bi', H, x) + g
        dual_inf = np.abs(grad + np.einsum('brv,br->bv', C, nu)).max(axis=1)
        prim_inf = np.abs(rc).max(axis=1)
        zsum = (np.abs(zl) + np.abs(zu)).sum(axis=1)
        s_d = np.maximum(o.smax, zsum / (2 * nr)) / o.smax
        s_c = s_d

        def compl(m_):
            return np.maximum(np.abs(zl * sl - m_[:, None]).max(axis=1), np.abs(zu * su - m_[:, None]).max(axis=1))

        E0 = np.maximum(np.maximum(dual_inf / s_d, prim_inf), compl(np.zeros(B)) / s_c)
        newly = (~done) & (E0 <= o.tol)
        status[newly] = STATUS_CONVERGED
        done |= newly
        bad = (~done) & ~np.isfinite(E0)
        status[bad] = STATUS_NUMERIC
        done |= bad
        if done.all() or it == o.max_iter:
            break
        iters[~done] += 1
        for _ in range(8):
            Emu = np.maximum(np.maximum(dual_inf / s_d, prim_inf), compl(mu) / s_c)
            red = (~done) & (Emu <= o.kappa_eps * mu) & (mu > mu_min)
            if not red.any():
                break
            mu = np.where(red, np.maximum(mu_min, np.minimum(o.kappa_mu * mu, mu ** o.theta_mu)), mu)
        with np.errstate(all='ignore'):     # finished instances are carried along unchanged
            isl, isu = 1.0 / sl, 1.0 / su
        sig = zl * isl + zu * isu
        nuhat = mu[:, None] * (isu - isl) + sig * rc
        Kk = H + np.einsum('bri,br,brj->bij', C, sig, C)
        rhs = -(grad + np.einsum('brv,br->bv', C, nuhat))
        dx = np.linalg.solve(Kk, rhs[:, :, None])[:, :, 0]
        ds = np.einsum('brv,bv->br', C, dx) + rc
        dzl = mu[:, None] * isl - zl - zl * isl * ds
        dzu = mu[:, None] * isu - zu + zu * isu * ds
        tau = np.maximum(o.tau_min, 1.0 - mu)[:, None]
        rp = np.maximum(-ds * isl, ds * isu).max(axis=1)
        rd = np.maximum(-dzl / zl, -dzu / zu).max(axis=1)
        ap = np.where(rp > tau[:, 0], tau[:, 0] / np.maximum(rp, 1e-300), 1.0)
        ad = np.where(rd > tau[:, 0], tau[:, 0] / np.maximum(rd, 1e-300), 1.0)
        act = ~done
        x = np.where(act[:, None], x + ap[:, None] * dx, x)
        s = np.where(act[:, None], s + ap[:, None] * ds, s)
        zl = np.where(act[:, None], zl + ad[:, None] * dzl, zl)
        zu = np.where(act[:, None], zu + ad[:, None] * dzu, zu)
    obj = 0.5 * np.einsum('bi,bij,bj->b', x, H, x) + np.einsum('bi,bi->b', g, x)
    return dict(x=x, obj=obj, status=status, iters=iters, zl=zl, zu=zu, s=s)


def kkt_residual(H, g, C, lo, hi, x, tol_act=1e-6):
    """Solver-independent optimality check: smallest achievable stationarity residual with multipliers of the right
    sign on the active rows (non-negative least squares), and the primal violation."""
    from scipy.optimize import nnls
    B = x.shape[0]
    stat = np.zeros(B)
    viol = np.zeros(B)
    for b in range(B):
        t = C[b] @ x[b]
        viol[b] = max(0.0, (lo[b] - t).max(), (t - hi[b]).max())
        grad = H[b] @ x[b] + g[b]
        scale = np.maximum(1.0, np.abs(hi[b] - lo[b]))
        cols = [C[b][r] for r in range(C.shape[1]) if t[r] >= hi[b][r] - tol_act * scale[r]]      # nu >= 0
        cols += [-C[b][r] for r in range(C.shape[1]) if t[r] <= lo[b][r] + tol_act * scale[r]]    # nu <= 0
        if cols:
            A = np.array(cols).T
            _, rn = nnls(A, -grad)
            stat[b] = rn
        else:
            stat[b] = np.linalg.norm(grad)
    return stat, viol
