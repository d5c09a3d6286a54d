"""The three reference NLPs as flat functions f(w, p), g(w, p) in the reference's own variable / constraint / parameter
layout.  TEST INFRASTRUCTURE ONLY.

``oracle.problems`` states the NLPs in stage form for the oracle's solver; this file states them the way the reference
hands them to ``ca.nlpsol`` so they can be compared value for value (objective, every constraint row, gradient and
Jacobian by complex step) with the graphs captured from the reference's source (``tests/golden/tapes``).  Batched over
leading axes and complex-safe.  w = [vec(X) column-major; vec(U)] (mpc_3d.py:69).
"""
import numpy as np

from . import models


def _split(w, nx, nu, N):
    nX = nx * (N + 1)
    X = w[..., :nX].reshape(w.shape[:-1] + (N + 1, nx))
    U = w[..., nX:nX + nu * N].reshape(w.shape[:-1] + (N, nu))
    return X, U


def pmpc_nlp(w, p, Qp, Qv, R, mu, N=15, Ts=0.002, g=-9.81):
    """mpc_3d.py:32-80.  p = [state(6); target(6)]; g = [X0 - p[:6]; X_{k+1} - f(X_k, U_k) for k < N]  (96 rows)."""
    p = np.broadcast_to(p, w.shape[:-1] + p.shape[-1:])
    X, U = _split(w, 6, 2, N)
    ref = p[..., None, 6:]
    e = X[..., :4] - ref[..., :4]
    wq = np.array([Qp, Qv, Qp, Qv])
    f = np.sum(wq * e[..., :N, :] ** 2, axis=(-1, -2)) + R * np.sum(U ** 2, axis=(-1, -2)) + np.sum(wq * e[..., N, :] ** 2, axis=-1)
    rows = [X[..., 0, :] - p[..., :6]]
    Fx = models.pmpc_step(X[..., :N, :], U, g, mu, Ts)
    rows.append((X[..., 1:, :] - Fx).reshape(w.shape[:-1] + (-1,)))
    return f, np.concatenate(rows, axis=-1)


def rmpc_nlp(w, p, N=20, Ts=0.002, Qp=80.0, Qv=2.0, Ru=0.02, Rdu=1.0, vmax=0.2, v_eps=0.1, gz=-9.81):
    """np_mpc_adaptive_with_linear_regressor.py:76-143.  p = [x0(4); u_prev(2); theta_hat(14); Rref((N+1)*4)];
    g = [X0 - x0; per stage: X_{k+1} - f (4), du_k (2), (vx - vmax, -vx - vmax, vy - vmax, -vy - vmax)]  (204 rows)."""
    p = np.broadcast_to(p, w.shape[:-1] + p.shape[-1:])
    X, U = _split(w, 4, 2, N)
    x0, u_prev, th = p[..., 0:4], p[..., 4:6], p[..., 6:20]
    ref = p[..., 20:].reshape(p.shape[:-1] + (N + 1, 4))
    dU = U - np.concatenate([u_prev[..., None, :], U[..., :-1, :]], axis=-2)
    wq = np.array([Qp, Qv, Qp, Qv])
    e = X - ref
    f = np.sum(wq * e ** 2, axis=(-1, -2)) + Ru * np.sum(U ** 2, axis=(-1, -2)) + Rdu * np.sum(dU ** 2, axis=(-1, -2))
    Fx = models.rmpc_step(X[..., :N, :], U, th[..., None, :], gz, v_eps, Ts)
    dyn = X[..., 1:, :] - Fx
    vx, vy = X[..., :N, 1], X[..., :N, 3]
    caps = np.stack([vx - vmax, -vx - vmax, vy - vmax, -vy - vmax], axis=-1)
    per_stage = np.concatenate([dyn, dU, caps], axis=-1).reshape(w.shape[:-1] + (-1,))
    return f, np.concatenate([X[..., 0, :] - x0, per_stage], axis=-1)


def rmpc_g_bounds(N=20, du_bounds=(-0.06, 0.06)):
    """lbg / ubg as the reference assembles them (:88-127)."""
    lo = [0.0] * 4
    hi = [0.0] * 4
    for _ in range(N):
        lo += [0.0] * 4 + [du_bounds[0]] * 2 + [-np.inf] * 4
        hi += [0.0] * 4 + [du_bounds[1]] * 2 + [0.0] * 4
    return np.array(lo), np.array(hi)


def lmpc_nlp(w, p, N=20, Ts=0.002, Q=(200.0, 2.0, 200.0, 2.0, 0.0, 0.0, 0.0, 0.0), Qt=(200.0, 2.0, 200.0, 2.0, 0.0, 0.0, 0.0, 0.0),
             R=(0.1, 0.1, 1.0, 1.0)):
    """rlmpc2.py:242-467.  p = [state(8); u_prev(2); pvec(34); target(8)] (:510); g = [X0 - state; X_{k+1} - f]  (168 rows)."""
    p = np.broadcast_to(p, w.shape[:-1] + p.shape[-1:])
    X, U = _split(w, 8, 2, N)
    state, u_prev, pvec, traj = p[..., 0:8], p[..., 8:10], p[..., 10:44], p[..., 44:52]
    dU = U - np.concatenate([u_prev[..., None, :], U[..., :-1, :]], axis=-2)
    e = X - traj[..., None, :]
    Q, Qt, R = np.asarray(Q), np.asarray(Qt), np.asarray(R)
    f = (np.sum(Q * e[..., :N, :] ** 2, axis=(-1, -2)) + np.sum(R[:2] * U ** 2, axis=(-1, -2))
         + np.sum(R[2:] * dU ** 2, axis=(-1, -2)) + np.sum(Qt * e[..., N, :] ** 2, axis=-1))
    Fx = models.lmpc_step(X[..., :N, :], U, pvec[..., None, :], Ts)
    dyn = (X[..., 1:, :] - Fx).reshape(w.shape[:-1] + (-1,))
    return f, np.concatenate([X[..., 0, :] - state, dyn], axis=-1)


def complex_step(fun, w, h=1e-30):
    """(grad f [..., n], Jacobian of g [..., m, n]) of fun: w -> (f, g) by the complex-step method."""
    n = w.shape[-1]
    wc = w[..., None, :] + 1j * h * np.eye(n)
    f, g = fun(wc)
    return f.imag / h, np.swapaxes(g.imag / h, -1, -2)
