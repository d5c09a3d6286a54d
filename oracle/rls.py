"""Recursive least squares with forgetting -- behaviour spec of the reference ``RLS`` class.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).
Follows RMPC/dev_dual/controller/np_mpc_adaptive_with_linear_regressor.py:10-30 operation by
operation (``np.outer(K, phi) @ P`` is NOT symmetrised) and the caller's regressor/measurement
construction in RMPC/dev_dual/rob_ctrl.py:335-343.
"""
import numpy as np


class RLS:
    def __init__(self, p, theta0=None, P0=1e3, lam=0.995):
        self.p = p
        self.theta = np.zeros(p) if theta0 is None else np.asarray(theta0, dtype=float).copy()
        self.P = np.eye(p) * float(P0)
        self.lam = float(lam)

    def update(self, phi, y):
        phi = np.asarray(phi, dtype=float).reshape(-1)
        y = float(np.asarray(y).reshape(()))
        denom = self.lam + phi @ self.P @ phi
        K = (self.P @ phi) / denom
        err = y - (phi @ self.theta)
        self.theta = self.theta + K * err
        self.P = (self.P - np.outer(K, phi) @ self.P) / self.lam

    def get(self):
        return self.theta.copy()


def rls_update_batch(theta, P, phi, y, lam=0.995):
    """Batched form of ``RLS.update``: theta [B,E,p], P [B,E,p,p], phi [B,p] (shared by the E estimators), y [B,E]."""
    theta = np.array(theta, dtype=np.float64, copy=True)
    P = np.array(P, dtype=np.float64, copy=True)
    B, E, p = theta.shape
    for b in range(B):
        for e in range(E):
            ph = phi[b]
            denom = lam + ph @ P[b, e] @ ph
            K = (P[b, e] @ ph) / denom
            err = y[b, e] - ph @ theta[b, e]
            theta[b, e] = theta[b, e] + K * err
            P[b, e] = (P[b, e] - np.outer(K, ph) @ P[b, e]) / lam
    return theta, P


def regressor(prev_state, v_eps):
    """phi of the *previous* state, rob_ctrl.py:338-339: [px, vx, py, vy, tanh(vx/eps), tanh(vy/eps), 1]."""
    ps = np.asarray(prev_state, dtype=np.float64)
    one = np.ones_like(ps[..., 0])
    return np.stack([ps[..., 0], ps[..., 1], ps[..., 2], ps[..., 3],
                     np.tanh(ps[..., 1] / v_eps), np.tanh(ps[..., 3] / v_eps), one], axis=-1)


def accel_measurement(xk, prev_state, Ts):
    """rob_ctrl.py:336-337: finite-difference acceleration (gravity term deliberately not removed)."""
    xk = np.asarray(xk, dtype=np.float64)
    ps = np.asarray(prev_state, dtype=np.float64)
    return np.stack([(xk[..., 1] - ps[..., 1]) / Ts, (xk[..., 3] - ps[..., 3]) / Ts], axis=-1)
