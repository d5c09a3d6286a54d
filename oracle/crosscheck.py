"""Third-party cross-check of the NLP optimum: condensed single shooting + scipy SLSQP.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  Shares only ``oracle.models`` /
``oracle.problems`` (the literal model and cost restatements) with ``oracle.ipm``; the
optimiser, its derivatives (finite differences inside scipy) and the formulation (states
eliminated by rollout) are unrelated.  Small instances only.
"""
import numpy as np
from scipy.optimize import minimize

from .problems import StageProblem


def _single(prob: StageProblem, b: int):
    """View of instance b as a B=1 problem (shares arrays)."""
    import copy
    p = copy.copy(prob)
    p.B = 1
    for name in ("x0", "wy", "ry", "wd", "wT", "rT"):
        setattr(p, name, getattr(prob, name)[b:b + 1])
    p.rows = [type(r)(r.ia, r.sa, r.ib, r.sb, r.lo[b:b + 1], r.hi[b:b + 1], r.skip0) for r in prob.rows]
    step = prob.step

    def step1(x, u):
        # evaluate through the batched step with this instance's parameters
        B = prob.B
        xx = np.repeat(x, B, axis=0)
        uu = np.repeat(u, B, axis=0)
        return step(xx, uu)[b:b + 1]

    p.step = step1
    return p


def solve_slsqp(prob: StageProblem, b: int = 0, U0=None, ftol=1e-15, maxiter=500):
    p = _single(prob, b)
    N, m, n = p.N, p.m, p.n

    def J(uflat):
        U = uflat.reshape(1, N, m)
        X = p.rollout(U)
        return float(p.objective(X, U)[0])

    cons = []
    bounds = [(None, None)] * (N * m)
    for r, row in enumerate(p.rows):
        if row.ib < 0 and row.ia >= n:
            j = row.ia - n
            for k in range(N):
                bounds[k * m + j] = (float(row.lo[0]) / row.sa, float(row.hi[0]) / row.sa)
        else:
            def cfun(uflat, r=r, row=row):
                U = uflat.reshape(1, N, m)
                X = p.rollout(U)
                t = p.row_values(X, U)[0, :, r]
                k0 = 1 if row.skip0 else 0
                return np.concatenate([t[k0:] - row.lo[0], row.hi[0] - t[k0:]])
            cons.append({"type": "ineq", "fun": cfun})
    u0 = np.zeros(N * m) if U0 is None else np.asarray(U0, dtype=np.float64).reshape(-1)
    res = minimize(J, u0, method="SLSQP", bounds=bounds, constraints=cons,
                   options={"ftol": ftol, "maxiter": maxiter})
    U = res.x.reshape(N, m)
    return dict(U=U, J=res.fun, success=res.success, nit=res.nit)
