"""Property tests of the device solver source (host harness): optimality, feasibility and invariances that hold for
any input, independent of the oracle's solver."""
import numpy as np
from hypothesis import given, settings, strategies as st

import dart_b200
from oracle import models, problems

N = 15


def _unpack(w):
    X = w[:, :(N + 1) * 6].reshape(-1, N + 1, 6)
    U = w[:, (N + 1) * 6:].reshape(-1, N, 2)
    return X, U


def _objective(X, U, target, Qp, Qv, R):
    """mpc_3d.py:40-65 summed literally."""
    J = 0.0
    for k in range(N + 1):
        J = J + Qp * ((X[:, k, 0] - target[:, 0]) ** 2 + (X[:, k, 2] - target[:, 2]) ** 2) \
              + Qv * ((X[:, k, 1] - target[:, 1]) ** 2 + (X[:, k, 3] - target[:, 3]) ** 2)
        if k < N:
            J = J + R * (U[:, k] ** 2).sum(axis=1)
    return J


inst = st.tuples(st.floats(-0.15, 0.15), st.floats(-0.2, 0.2), st.floats(-0.1, 0.1), st.floats(-0.2, 0.2),
                 st.floats(-0.125, 0.125), st.floats(-0.125, 0.125), st.sampled_from([(600.0, 5.0, 0.1), (400.0, 2.5, 0.2), (200.0, 2.0, 0.2)]),
                 st.sampled_from([0.05, 0.1, 0.2]))


@settings(max_examples=40, deadline=None)
@given(st.lists(inst, min_size=1, max_size=6), st.integers(0, 2 ** 31 - 1))
def test_pmpc_solution_is_feasible_and_locally_optimal(hostemu, cases, seed):
    B = len(cases)
    state = np.array([[c[0], c[1], c[2], c[3], 0.43, 0.0] for c in cases])
    target = np.array([[c[4], 0.0, c[5], 0.0, 0.4, 0.0] for c in cases])
    aux = np.array([[c[6][0], c[6][1], c[6][2], c[7]] for c in cases])
    out = hostemu.solve(dart_b200.pmpc_cfg(), state, target, aux)
    assert (out["status"] == 0).all()
    X, U = _unpack(out["w"])
    # bounds, initial condition, dynamics of the full 6-state model (reference's equality constraints)
    assert (np.abs(U) <= 0.6 + 1e-12).all() and np.array_equal(X[:, 0], state)
    for k in range(N):
        assert np.abs(models.pmpc_step(X[:, k], U[:, k], -9.81, aux[:, 3], 0.002) - X[:, k + 1]).max() < 1e-8
    # the returned loss is the reference's objective at the returned point
    J = _objective(X, U, target, aux[:, 0], aux[:, 1], aux[:, 2])
    assert np.abs(J - out["J"]).max() <= 1e-9 * max(1.0, np.abs(J).max())
    # local optimality: no feasible perturbation of the controls (rolled out exactly) does better
    p = problems.pmpc_problem(state, target, Qp=aux[:, 0], Qv=aux[:, 1], R=aux[:, 2], mu=aux[:, 3])
    rng = np.random.default_rng(seed)
    for scale in (1e-1, 1e-2, 1e-3):
        Up = np.clip(U + scale * rng.standard_normal(U.shape), -0.6, 0.6)
        Jp = p.objective(p.rollout(Up), Up)
        assert (Jp >= out["J"] - 1e-7 * np.maximum(1.0, np.abs(out["J"]))).all()


@settings(max_examples=15, deadline=None)
@given(st.lists(inst, min_size=2, max_size=8), st.integers(0, 2 ** 31 - 1))
def test_batch_order_and_neighbours_do_not_matter(hostemu, cases, seed):
    state = np.array([[c[0], c[1], c[2], c[3], 0.43, 0.0] for c in cases])
    target = np.array([[c[4], 0.0, c[5], 0.0, 0.4, 0.0] for c in cases])
    aux = np.array([[c[6][0], c[6][1], c[6][2], c[7]] for c in cases])
    cfg = dart_b200.pmpc_cfg()
    a = hostemu.solve(cfg, state, target, aux)
    perm = np.random.default_rng(seed).permutation(len(cases))
    b = hostemu.solve(cfg, state[perm], target[perm], aux[perm])
    assert np.array_equal(a["u0"][perm], b["u0"]) and np.array_equal(a["J"][perm], b["J"])
    one = hostemu.solve(cfg, state[:1], target[:1], aux[:1])
    assert np.array_equal(one["u0"][0], a["u0"][0]) and one["J"][0] == a["J"][0]


def test_quaternion_epilogue_formula():
    u = np.array([[0.3, -0.2], [0.0, 0.0], [-0.6, 0.6]])
    from scipy.spatial.transform import Rotation as Rot
    q = models.tilt_to_quat(u)
    for i in range(3):
        ref = Rot.from_euler("xyz", [u[i, 1], -u[i, 0], 0.0]).as_quat()   # xyzw (rob_ctrl.py:355 uses scalar_first=True)
        assert np.allclose(q[i], [ref[3], ref[0], ref[1], ref[2]], atol=1e-15)
