"""The three reference NLPs in stage form, batched over independent instances.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).

A ``StageProblem`` is
    min  sum_{k<N} [ sum_i wy_i (y_ki - ry_ki)^2 + sum_j wd_j (u_kj - x_k[n-m+j])^2 ] + sum_i wT_i (x_Ni - rT_i)^2
    s.t. x_0 given, x_{k+1} = step(x_k, u_k),  lo_r <= c_r . y_k <= hi_r  for every row r and stage k
with y_k = [x_k; u_k].  Where the reference penalises/limits the tilt *rate*
(u_k - u_{k-1}) the previous control is carried as ``m`` extra trailing states
(x_{k+1}[n-m+j] = u_kj, x_0[n-m+j] = u_prev_j), which is algebraically the same
NLP as the reference's.  Objective values are identical to the reference's
``sol['f']`` by construction (``objective``).
"""
from dataclasses import dataclass, field
from typing import Callable, List

import numpy as np

from . import models


@dataclass
class Row:
    ia: int            # index into y = [x; u]
    sa: float
    ib: int            # -1: single-entry row
    sb: float
    lo: np.ndarray     # [B]
    hi: np.ndarray     # [B]
    skip0: bool = False  # row acts on fixed quantities at k = 0 -> not a constraint on decision variables


@dataclass
class StageProblem:
    name: str
    B: int
    n: int
    m: int
    N: int
    naug: int
    x0: np.ndarray                 # [B, n]
    wy: np.ndarray                 # [B, n+m]
    ry: np.ndarray                 # [B, N, n+m]
    wd: np.ndarray                 # [B, m]  (zeros when naug == 0)
    wT: np.ndarray                 # [B, n]
    rT: np.ndarray                 # [B, n]
    rows: List[Row]
    step: Callable                 # (x[B,...,n], u[B,...,m]) -> [B,...,n]
    nphys: int = field(default=0)  # physical state count (n - naug)
    aux: dict = field(default_factory=dict)

    def __post_init__(self):
        self.nphys = self.n - self.naug

    # ---- objective exactly as the reference sums it
    def objective(self, X, U):
        n, m, N = self.n, self.m, self.N
        y = np.concatenate([X[:, :N], U], axis=-1)
        J = np.sum(self.wy[:, None, :] * (y - self.ry) ** 2, axis=(1, 2))
        if self.naug:
            e = U - X[:, :N, n - m:]
            J = J + np.sum(self.wd[:, None, :] * e ** 2, axis=(1, 2))
        J = J + np.sum(self.wT * (X[:, N] - self.rT) ** 2, axis=1)
        return J

    def rollout(self, U):
        X = np.empty((self.B, self.N + 1, self.n))
        X[:, 0] = self.x0
        for k in range(self.N):
            X[:, k + 1] = self.step(X[:, k], U[:, k])
        return X

    def row_values(self, X, U):
        y = np.concatenate([X[:, :self.N], U], axis=-1)
        t = np.empty((self.B, self.N, len(self.rows)))
        for r, row in enumerate(self.rows):
            t[:, :, r] = row.sa * y[:, :, row.ia] + (row.sb * y[:, :, row.ib] if row.ib >= 0 else 0.0)
        return t

    def row_mask(self):
        msk = np.ones((self.N, len(self.rows)))
        for r, row in enumerate(self.rows):
            if row.skip0:
                msk[0, r] = 0.0
        return msk


def _bc(p, x):
    """Reshape a per-instance parameter [B] / [B, q] so it broadcasts against x[B, ..., n]."""
    p = np.asarray(p, dtype=np.float64)
    extra = x.ndim - 2
    if p.ndim == 1:
        return p.reshape((p.shape[0],) + (1,) * extra)
    return p.reshape((p.shape[0],) + (1,) * extra + (p.shape[1],))


def _col(v, B):
    v = np.asarray(v, dtype=np.float64)
    return np.broadcast_to(v, (B,)).copy() if v.ndim == 0 else v.copy()


# ----------------------------------------------------------------------------- PMPC
def pmpc_problem(state, target, Ts=0.002, N=15, Qp=400.0, Qv=2.0, R=0.2, mu=0.1,
                 u_bounds=(-0.6, 0.6), g=-9.81):
    """NLP of ``PMPC.__init__`` (mpc_3d.py:28-85) with p = [state; target] as in ``solve`` (:115-138).

    state, target: [B, 6].  Qp/Qv/R/mu may be per-instance arrays [B].

    The z rows (pz, vz) carry no cost (mpc_3d.py:44-46,62-64 index only 0..3), no bound, and
    feed nothing back into x/y, so they are an always-feasible appendix of the NLP whose
    multipliers vanish at any KKT point.  The oracle optimises the 4-state x/y NLP and obtains
    the z columns of ``X`` by rolling ``pmpc_step`` forward (``pmpc_full_states``); the KKT
    check in ``tests/test_oracle.py`` is done on the full 6-state NLP.
    """
    state = np.atleast_2d(np.asarray(state, dtype=np.float64))
    target = np.atleast_2d(np.asarray(target, dtype=np.float64))
    B = state.shape[0]
    Qp, Qv, R, mu = (_col(v, B) for v in (Qp, Qv, R, mu))
    wy = np.stack([Qp, Qv, Qp, Qv, R, R], axis=1)
    ry = np.zeros((B, N, 6))
    ry[:, :, :4] = target[:, None, :4]
    wT = np.stack([Qp, Qv, Qp, Qv], axis=1)
    lo, hi = _col(u_bounds[0], B), _col(u_bounds[1], B)
    rows = [Row(4, 1.0, -1, 0.0, lo, hi), Row(5, 1.0, -1, 0.0, lo, hi)]

    def step(x, u):
        x6 = np.concatenate([x, np.zeros(x.shape[:-1] + (2,), dtype=x.dtype)], axis=-1)
        return models.pmpc_step(x6, u, g, _bc(mu, x), Ts)[..., :4]

    prob = StageProblem("pmpc", B, 4, 2, N, 0, state[:, :4].copy(), wy, ry, np.zeros((B, 2)), wT,
                        target[:, :4].copy(), rows, step)
    prob.aux = dict(state6=state.copy(), g=g, mu=mu, Ts=Ts)
    return prob


def pmpc_full_states(prob, U):
    """X [B, N+1, 6] of the reference's decision vector for controls U: roll ``pmpc_step`` from state6."""
    a = prob.aux
    X = np.empty((prob.B, prob.N + 1, 6))
    X[:, 0] = a["state6"]
    for k in range(prob.N):
        X[:, k + 1] = models.pmpc_step(X[:, k], U[:, k], a["g"], a["mu"], a["Ts"])
    return X


# ----------------------------------------------------------------------------- RMPC
def rmpc_problem(x0, u_prev, theta_hat, Rref_flat, Ts=0.002, N=20, Qp=80.0, Qv=2.0, Ru=0.02,
                 Rdu=1.0, u_bounds=(-0.6, 0.6), du_bounds=(-0.06, 0.06), vmax=0.2, v_eps=0.1,
                 gz=-9.81):
    """NLP of ``AdaptiveNPMPCSmooth.__init__`` (np_mpc_adaptive_with_linear_regressor.py:65-168).

    x0 [B,4], u_prev [B,2], theta_hat [B,14], Rref_flat [B,(N+1)*4].  The four one-sided
    velocity caps per stage (:124-127) are stated as two two-sided rows; at k = 0 they act on
    the fixed x_0 (skip0).
    """
    x0 = np.atleast_2d(np.asarray(x0, dtype=np.float64))
    B = x0.shape[0]
    u_prev = np.atleast_2d(np.asarray(u_prev, dtype=np.float64))
    th = np.atleast_2d(np.asarray(theta_hat, dtype=np.float64))
    ref = np.atleast_2d(np.asarray(Rref_flat, dtype=np.float64)).reshape(B, N + 1, 4)
    Qp, Qv, Ru, Rdu = (_col(v, B) for v in (Qp, Qv, Ru, Rdu))
    z = np.zeros(B)
    wy = np.stack([Qp, Qv, Qp, Qv, z, z, Ru, Ru], axis=1)
    ry = np.zeros((B, N, 8))
    ry[:, :, :4] = ref[:, :N]
    wd = np.stack([Rdu, Rdu], axis=1)
    wT = np.stack([Qp, Qv, Qp, Qv, z, z], axis=1)
    rT = np.zeros((B, 6))
    rT[:, :4] = ref[:, N]
    ulo, uhi = _col(u_bounds[0], B), _col(u_bounds[1], B)
    dlo, dhi = _col(du_bounds[0], B), _col(du_bounds[1], B)
    vm = _col(vmax, B)
    rows = [Row(6, 1.0, -1, 0.0, ulo, uhi), Row(7, 1.0, -1, 0.0, ulo, uhi),
            Row(6, 1.0, 4, -1.0, dlo, dhi), Row(7, 1.0, 5, -1.0, dlo, dhi),
            Row(1, 1.0, -1, 0.0, -vm, vm, skip0=True), Row(3, 1.0, -1, 0.0, -vm, vm, skip0=True)]

    def step(x, u):
        xn = models.rmpc_step(x[..., :4], u, _bc(th, x), gz, v_eps, Ts)
        return np.concatenate([xn, u + 0 * xn[..., :2]], axis=-1)

    return StageProblem("rmpc", B, 6, 2, N, 2, np.concatenate([x0, u_prev], axis=1), wy, ry, wd,
                        wT, rT, rows, step)


def build_ref_traj(x_now, r_v, target, N, nx, step_fraction=0.2):
    """``AdaptiveNPMPCSmooth.build_ref_traj`` (np_mpc_adaptive_with_linear_regressor.py:201-210), batched."""
    r_v = np.atleast_2d(np.asarray(r_v, dtype=np.float64))
    target = np.atleast_2d(np.asarray(target, dtype=np.float64))
    B = r_v.shape[0]
    Rr = np.zeros((B, N + 1, nx))
    for i in range(N + 1):
        w = 1.0 - (1.0 - step_fraction) ** (i + 1)
        r_i = r_v + w * (target - r_v)
        Rr[:, i, 0] = r_i[:, 0]
        Rr[:, i, 2] = r_i[:, 2]
    return Rr.reshape(B, -1)


def reference_governor(r_v, target, dr_max=0.01, alpha_rg=0.5):
    """rob_ctrl.py:346-348, batched: r_v += alpha_rg * clip(target - r_v, +-dr_max) on the two positions."""
    r_v = np.array(r_v, dtype=np.float64, copy=True)
    target = np.asarray(target, dtype=np.float64)
    for i in (0, 2):
        r_v[..., i] = r_v[..., i] + alpha_rg * np.clip(target[..., i] - r_v[..., i], -dr_max, dr_max)
    return r_v


# ----------------------------------------------------------------------------- LMPC
def lmpc_problem(state, u_prev, pvec, target, Ts=0.002, N=20,
                 Q=(200.0, 2.0, 200.0, 2.0, 0.0, 0.0, 0.0, 0.0),
                 Qt=(200.0, 2.0, 200.0, 2.0, 0.0, 0.0, 0.0, 0.0),
                 R=(0.1, 0.1, 1.0, 1.0), u_bounds=(-0.4, 0.4)):
    """NLP of ``RLMPC._solver_worker`` (rlmpc2.py:236-491), p = [state(8); control(2); pvec(34); target(8)] (:510)."""
    state = np.atleast_2d(np.asarray(state, dtype=np.float64))
    B = state.shape[0]
    u_prev = np.atleast_2d(np.asarray(u_prev, dtype=np.float64))
    pvec = np.atleast_2d(np.asarray(pvec, dtype=np.float64))
    target = np.atleast_2d(np.asarray(target, dtype=np.float64))
    Q = np.broadcast_to(np.asarray(Q, dtype=np.float64), (B, 8))
    Qt = np.broadcast_to(np.asarray(Qt, dtype=np.float64), (B, 8))
    R = np.broadcast_to(np.asarray(R, dtype=np.float64), (B, 4))
    wy = np.concatenate([Q, np.zeros((B, 2)), R[:, 0:2]], axis=1)
    ry = np.zeros((B, N, 12))
    ry[:, :, :8] = target[:, None, :]
    wd = R[:, 2:4].copy()
    wT = np.concatenate([Qt, np.zeros((B, 2))], axis=1)
    rT = np.concatenate([target, np.zeros((B, 2))], axis=1)
    lo, hi = _col(u_bounds[0], B), _col(u_bounds[1], B)
    rows = [Row(10, 1.0, -1, 0.0, lo, hi), Row(11, 1.0, -1, 0.0, lo, hi)]

    def step(x, u):
        xn = models.lmpc_step(x[..., :8], u, _bc(pvec, x), Ts)
        return np.concatenate([xn, u + 0 * xn[..., :2]], axis=-1)

    return StageProblem("lmpc", B, 10, 2, N, 2, np.concatenate([state, u_prev], axis=1), wy, ry, wd,
                        wT, rT, rows, step)


# ----------------------------------------------------------------------------- reference decision-vector layout
def pack_w(prob: StageProblem, X, U):
    """w = [vec(X) column-major ; vec(U)] with physical states only (mpc_3d.py:69,137)."""
    return np.concatenate([X[:, :, :prob.nphys].reshape(prob.B, -1), U.reshape(prob.B, -1)], axis=1)


def unpack_w(prob: StageProblem, w):
    w = np.atleast_2d(np.asarray(w, dtype=np.float64))
    nX = (prob.N + 1) * prob.nphys
    Xp = w[:, :nX].reshape(prob.B, prob.N + 1, prob.nphys)
    U = w[:, nX:].reshape(prob.B, prob.N, prob.m)
    X = np.zeros((prob.B, prob.N + 1, prob.n))
    X[:, :, :prob.nphys] = Xp
    if prob.naug:
        X[:, 0, prob.nphys:] = prob.x0[:, prob.nphys:]
        X[:, 1:, prob.nphys:] = U
    return X, U
