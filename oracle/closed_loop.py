"""CPU restatement of the surrogate closed loop (plant step + logger metrics) for checking the device episodes.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  Plant: the PMPC model (mpc_3d.py:87-104) with per-instance viscous
mu and an optional Coulomb term; metrics as PMPC/src/logger.py:155-176.
"""
import numpy as np

from . import ipm, problems


def plant_step(x, u, mu, coulomb, Ts=0.002, g=-9.81):
    def f(x):
        vn = -g * (u[:, 0] ** 2 + u[:, 1] ** 2)
        ax = g * np.sin(u[:, 0]) - mu * x[:, 1] - coulomb * abs(g) * np.tanh(x[:, 1] / 0.01)
        ay = g * np.sin(u[:, 1]) - mu * x[:, 3] - coulomb * abs(g) * np.tanh(x[:, 3] / 0.01)
        return np.stack([x[:, 1], ax, x[:, 3], ay, vn, (vn - x[:, 5]) / Ts], axis=1)
    k1 = f(x); k2 = f(x + Ts / 2 * k1); k3 = f(x + Ts / 2 * k2); k4 = f(x + Ts * k3)
    return x + Ts / 6 * (k1 + 2 * k2 + 2 * k3 + k4)


def pmpc_episode(state, target, params, steps, mu_plant=None, coulomb=None, tol=0.01, Ts=0.002):
    B = state.shape[0]
    x = state.copy()
    mu_plant = params[:, 3] if mu_plant is None else mu_plant
    coulomb = np.zeros(B) if coulomb is None else coulomb
    conv = np.full(B, -1.0); effort = np.zeros(B); us = []
    for k in range(steps):
        r = ipm.solve(problems.pmpc_problem(x, target, Qp=params[:, 0], Qv=params[:, 1], R=params[:, 2], mu=params[:, 3], Ts=Ts))
        u = r["U"][:, 0]
        e = np.linalg.norm(x[:, [0, 2]] - target[:, [0, 2]], axis=1)
        conv = np.where((conv < 0) & (e < tol), k * Ts, conv)
        effort += np.linalg.norm(u, axis=1) * Ts
        x = plant_step(x, u, mu_plant, coulomb, Ts)
        us.append(u)
    return dict(state=x, u=np.array(us), conv_time=conv, effort=effort)
