"""BASELINE.json's full sizes on the GPU, checked through size-independent properties (the oracle *solver* is too slow
there; the oracle's literal model/objective restatements are vectorised and cheap): every instance converged, the
returned decision vector satisfies the reference's equality constraints and bounds, the returned loss is the
reference's objective at that point, and a sample of instances matches the oracle solver."""
import numpy as np
import pytest

import dart_b200
from oracle import ipm, models, problems

pytestmark = pytest.mark.gpu
W = dart_b200.workloads


def test_pmpc_131k_instances(built):
    c, aux = W.pmpc_inputs(7282)                       # 18 x 7282 = 131 076 instances (config 2 tiled up)
    B, N = aux.shape[0], 15
    out = dart_b200.NMPCEngine(dart_b200.pmpc_cfg(), device=0).solve(c["state"], c["target"], aux=aux)
    assert (out["status"] == 0).all() and out["iters"].max() <= 25
    X = out["w"][:, :(N + 1) * 6].reshape(B, N + 1, 6); U = out["w"][:, (N + 1) * 6:].reshape(B, N, 2)
    assert np.abs(U).max() <= 0.6 + 1e-12 and np.array_equal(X[:, 0], c["state"])
    worst = 0.0
    J = np.zeros(B)
    for k in range(N + 1):
        if k < N:
            worst = max(worst, np.abs(models.pmpc_step(X[:, k], U[:, k], -9.81, c["mu"], 0.002) - X[:, k + 1]).max())
            J += c["R"] * (U[:, k] ** 2).sum(axis=1)
        J += c["Qp"] * ((X[:, k, 0] - c["target"][:, 0]) ** 2 + (X[:, k, 2] - c["target"][:, 2]) ** 2)
        J += c["Qv"] * ((X[:, k, 1] - c["target"][:, 1]) ** 2 + (X[:, k, 3] - c["target"][:, 3]) ** 2)
    assert worst < 1e-8
    assert (np.abs(J - out["J"]) <= 1e-9 * np.maximum(1.0, np.abs(J))).all()
    idx = np.random.default_rng(0).choice(B, 96, replace=False)
    ref = ipm.solve(problems.pmpc_problem(c["state"][idx], c["target"][idx], Qp=c["Qp"][idx], Qv=c["Qv"][idx], R=c["R"][idx], mu=c["mu"][idx]))
    assert np.abs(out["u0"][idx] - ref["U"][:, 0]).max() <= 1e-4
    assert (np.abs(out["J"][idx] - ref["J"]) / np.abs(ref["J"])).max() <= 1e-6


def test_rmpc_config3_size(built):
    d = W.rmpc_inputs(4096)
    B, N = 4096, 20
    out = dart_b200.NMPCEngine(dart_b200.rmpc_cfg(), device=0).solve(d["x0"], d["ref"], aux=d["aux"])
    assert (out["status"] == 0).all()
    X = out["w"][:, :(N + 1) * 4].reshape(B, N + 1, 4); U = out["w"][:, (N + 1) * 4:].reshape(B, N, 2)
    assert np.abs(U).max() <= 0.6 + 1e-12
    dU = np.diff(np.concatenate([d["u_prev"][:, None, :], U], axis=1), axis=1)
    assert np.abs(dU).max() <= 0.06 + 1e-9                                  # tilt-rate rows (np_mpc...:117-122)
    assert np.abs(X[:, 1:N, [1, 3]]).max() <= 0.2 + 1e-9                    # velocity caps on stages 1..N-1 (:124-127)
    for k in range(N):
        assert np.abs(models.rmpc_step(X[:, k], U[:, k], d["theta"], -9.81, 0.1, 0.002) - X[:, k + 1]).max() < 1e-8
    p = problems.rmpc_problem(d["x0"], d["u_prev"], d["theta"], d["ref"])
    Xa = np.concatenate([X, np.concatenate([d["u_prev"][:, None, :], U], axis=1)], axis=2)
    assert (np.abs(p.objective(Xa, U) - out["J"]) <= 1e-9 * np.maximum(1.0, np.abs(out["J"]))).all()


def test_lmpc_config4_size(built):
    d = W.lmpc_inputs(16384)
    B, N = 16384, 20
    out = dart_b200.NMPCEngine(dart_b200.lmpc_cfg(), device=0).solve(d["x0"], d["ref"], aux=d["aux"])
    assert (out["status"] == 0).all()
    X = out["w"][:, :(N + 1) * 8].reshape(B, N + 1, 8); U = out["w"][:, (N + 1) * 8:].reshape(B, N, 2)
    assert np.abs(U).max() <= 0.4 + 1e-12 and np.array_equal(X[:, 0], d["x0"])
    for k in range(N):
        assert np.abs(models.lmpc_step(X[:, k], U[:, k], d["pvec"], 0.002) - X[:, k + 1]).max() < 1e-8
    p = problems.lmpc_problem(d["x0"], d["u_prev"], d["pvec"], d["ref"])
    Xa = np.concatenate([X, np.concatenate([d["u_prev"][:, None, :], U], axis=1)], axis=2)
    assert (np.abs(p.objective(Xa, U) - out["J"]) <= 1e-9 * np.maximum(1.0, np.abs(out["J"]))).all()
