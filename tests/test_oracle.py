"""The oracle against itself: two unrelated solvers, KKT conditions of the reference-layout NLP, anchors."""
import numpy as np
import pytest

from oracle import crosscheck, ipm, models, problems, rls
from tests import helpers


def test_pmpc_anchor_values():
    # sanity anchors recorded in SURVEY.md 8(c) (condensed SLSQP/L-BFGS-B probes, cube weights, mu = 0.1)
    p = problems.pmpc_problem([[0, 0, 0, 0, .43, 0], [.098, .05, .049, .02, .43, 0]], [[.1, 0, .05, 0, .4, 0]] * 2,
                              Qp=600, Qv=5, R=0.1, mu=0.1)
    r = ipm.solve(p)
    assert (r["status"] == 0).all()
    assert abs(r["J"][0] - 118.947790217) < 1e-6 and np.allclose(r["U"][0, 0], [-0.6, -0.6], atol=1e-6)
    assert abs(r["J"][1] - 0.140159134621) < 1e-9 and np.allclose(r["U"][1, 0], [0.291456, 0.116327], atol=1e-6)


def test_rmpc_anchor_values():
    rv = np.array([[.005, 0, .005, 0]]); tg = np.array([[.05, 0, .05, 0]])
    ref = problems.build_ref_traj(None, rv, tg, 20, 4, 0.2)
    r = ipm.solve(problems.rmpc_problem(np.zeros((1, 4)), np.zeros((1, 2)), np.zeros((1, 14)), ref))
    assert r["status"][0] == 0 and abs(r["J"][0] - 6.087598328) < 1e-8
    assert np.allclose(r["U"][0, 0], [-0.0502635, -0.0502635], atol=1e-6)


@pytest.mark.parametrize("which", ["pmpc", "rmpc", "lmpc"])
def test_ipm_agrees_with_slsqp(which):
    if which == "pmpc":
        _, _, p = helpers.pmpc_case(1)
        idx = [0, 7, 13]
    elif which == "rmpc":
        _, p = helpers.rmpc_case(2)
        idx = [0]
    else:
        _, p = helpers.lmpc_case(2)
        idx = [1]
    r = ipm.solve(p)
    assert (r["status"] == 0).all()
    for b in idx:
        c = crosscheck.solve_slsqp(p, b)
        assert c["success"]
        assert abs(c["J"] - r["J"][b]) <= 1e-6 * max(1.0, abs(r["J"][b]))
        assert np.abs(c["U"][0] - r["U"][b, 0]).max() <= 1e-4


def test_pmpc_full_nlp_kkt():
    """KKT of the reference's 6-state NLP (z rows included, multipliers zero) at the 4-state optimum."""
    c, _, p = helpers.pmpc_case(1)
    r = ipm.solve(p)
    X6 = problems.pmpc_full_states(p, r["U"])
    a = p.aux
    # dynamics feasibility of every row incl. z
    for k in range(p.N):
        nxt = models.pmpc_step(X6[:, k], r["U"][:, k], a["g"], a["mu"], a["Ts"])
        assert np.abs(nxt - X6[:, k + 1]).max() < 1e-9
    assert np.abs(X6[:, :, :4] - r["X"]).max() < 1e-7
    # objective as the reference sums it (mpc_3d.py:40-65) equals the reduced objective
    tgt = c["target"]
    J = np.zeros(p.B)
    for k in range(p.N + 1):
        pos = (X6[:, k, 0] - tgt[:, 0]) ** 2 + (X6[:, k, 2] - tgt[:, 2]) ** 2
        vel = (X6[:, k, 1] - tgt[:, 1]) ** 2 + (X6[:, k, 3] - tgt[:, 3]) ** 2
        J += c["Qp"] * pos + c["Qv"] * vel
        if k < p.N:
            J += c["R"] * (r["U"][:, k] ** 2).sum(axis=1)
    assert np.abs(J - r["J"]).max() <= 1e-7 * np.abs(J).max()
    assert r["kkt"].max() <= 1e-8


def test_jacobians_complex_step_vs_fd():
    _, p = helpers.lmpc_case(2)
    rng = np.random.default_rng(0)
    X = p.x0[:, None, :] + 0.01 * rng.standard_normal((p.B, p.N + 1, p.n))
    U = 0.1 * rng.standard_normal((p.B, p.N, p.m))
    F, A, Bm = ipm.jacobians(p, X, U)
    h = 1e-6
    for j in range(p.n):
        Xp = X.copy(); Xm = X.copy(); Xp[:, :, j] += h; Xm[:, :, j] -= h
        fd = (p.step(Xp[:, :p.N], U) - p.step(Xm[:, :p.N], U)) / (2 * h)
        assert np.abs(fd - A[..., j]).max() < 1e-6


def test_rls_batch_equals_class():
    rng = np.random.default_rng(1)
    B = 3
    est = [[rls.RLS(7), rls.RLS(7)] for _ in range(B)]
    theta = np.zeros((B, 2, 7)); P = np.tile(np.eye(7) * 1e3, (B, 2, 1, 1))
    for _ in range(50):
        phi = rng.standard_normal((B, 7)); y = rng.standard_normal((B, 2))
        theta, P = rls.rls_update_batch(theta, P, phi, y)
        for b in range(B):
            for e in range(2):
                est[b][e].update(phi[b], y[b, e])
    for b in range(B):
        for e in range(2):
            assert np.array_equal(theta[b, e], est[b][e].get())
            assert np.array_equal(P[b, e], est[b][e].P)


def test_build_ref_traj_and_governor():
    rv = np.array([[0.0, 0, 0.0, 0]]); tg = np.array([[0.1, 0, -0.05, 0]])
    R = problems.build_ref_traj(None, rv, tg, 20, 4, 0.2).reshape(21, 4)
    assert np.allclose(R[0], [0.1 * 0.2, 0, -0.05 * 0.2, 0]) and np.allclose(R[:, [1, 3]], 0)
    assert np.allclose(R[20, 0], 0.1 * (1 - 0.8 ** 21))
    r2 = problems.reference_governor(rv, tg)
    assert np.allclose(r2, [[0.005, 0, -0.005, 0]])
