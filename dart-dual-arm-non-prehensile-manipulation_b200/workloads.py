"""Seeded synthetic inputs for the BASELINE.json configurations.

Pure numpy; used by bench.py, the tests and smoke() so the CUDA path, the oracle and the CPU
baseline all see identical inputs.  Nothing here computes MPC results.
"""
import numpy as np

# PMPC/main_parallel.py:107-118 -- per-shape cost weights (Qp, Qv, R)
PMPC_SHAPE_WEIGHTS = {"cube": (600.0, 5.0, 0.1), "cylinder": (400.0, 2.5, 0.2), "sphere": (200.0, 2.0, 0.2)}
PMPC_MASSES = (1.0, 2.0)            # NLP-inert (mass never enters mpc_3d.py)
PMPC_FRICTIONS = (0.05, 0.10, 0.20)


def pmpc_objects():
    """The 18 shape x mass x friction objects: list of dict(shape, mass, mu, Qp, Qv, R)."""
    objs = []
    for shape, (Qp, Qv, R) in PMPC_SHAPE_WEIGHTS.items():
        for mass in PMPC_MASSES:
            for mu in PMPC_FRICTIONS:
                objs.append(dict(shape=shape, mass=mass, mu=mu, Qp=Qp, Qv=Qv, R=R))
    return objs


def pmpc_config1():
    """BASELINE config 1: cube, 1 kg, mu=0.10, PMPC/main.py:59-69 weights, README example target."""
    return dict(state=np.array([[0.0, 0.0, 0.0, 0.0, 0.43, 0.0]]),
                target=np.array([[0.1, 0.0, 0.05, 0.0, 0.4, 0.0]]),
                Qp=np.array([400.0]), Qv=np.array([2.0]), R=np.array([0.2]), mu=np.array([0.10]))


def pmpc_config2(states_per_object=64, seed=1):
    """BASELINE config 2: 18 objects x ``states_per_object`` random (x0, target) pairs. B = 18*S."""
    rng = np.random.default_rng(seed)
    objs = pmpc_objects()
    S = states_per_object
    B = len(objs) * S
    state = np.zeros((B, 6))
    target = np.zeros((B, 6))
    state[:, 0] = rng.uniform(-0.15, 0.15, B)
    state[:, 2] = rng.uniform(-0.10, 0.10, B)
    state[:, 1] = rng.uniform(-0.2, 0.2, B)
    state[:, 3] = rng.uniform(-0.2, 0.2, B)
    state[:, 4] = 0.43
    target[:, 0] = rng.uniform(-0.125, 0.125, B)
    target[:, 2] = rng.uniform(-0.125, 0.125, B)
    target[:, 4] = 0.4
    rep = lambda key: np.repeat(np.array([o[key] for o in objs], dtype=np.float64), S)
    return dict(state=state, target=target, Qp=rep("Qp"), Qv=rep("Qv"), R=rep("R"), mu=rep("mu"),
                mass=rep("mass"), shape=np.repeat(np.arange(len(objs)) // 6, S))


def rmpc_config3(B=4096, seed=2):
    """BASELINE config 3 initial conditions: x0, target offsets, surrogate-plant friction/damping."""
    rng = np.random.default_rng(seed)
    x0 = np.zeros((B, 4))
    x0[:, 0] = rng.uniform(-0.1, 0.1, B)
    x0[:, 2] = rng.uniform(-0.1, 0.1, B)
    x0[:, 1] = rng.uniform(-0.15, 0.15, B)
    x0[:, 3] = rng.uniform(-0.15, 0.15, B)
    target = np.zeros((B, 4))
    target[:, 0] = rng.uniform(-0.1, 0.1, B)
    target[:, 2] = rng.uniform(-0.1, 0.1, B)
    mu_plant = rng.choice(np.array([0.05, 0.1, 0.2]), B)
    c_plant = rng.uniform(0.0, 0.5, B)
    return dict(x0=x0, target=target, mu_plant=mu_plant, c_plant=c_plant)


def rmpc_plant_step(x, u, mu_p, c_p, Ts=0.002, gz=-9.81):
    """Surrogate plant of SURVEY 8(d) config 3: v' = gz sin u - mu g tanh(v/.01) - c v, four explicit Euler sub-steps."""
    x = np.array(x, dtype=np.float64, copy=True)
    for _ in range(4):
        h = Ts / 4
        ax = gz * np.sin(u[:, 0]) - mu_p * 9.81 * np.tanh(x[:, 1] / 0.01) - c_p * x[:, 1]
        ay = gz * np.sin(u[:, 1]) - mu_p * 9.81 * np.tanh(x[:, 3] / 0.01) - c_p * x[:, 3]
        x[:, 0] += h * x[:, 1]; x[:, 2] += h * x[:, 3]
        x[:, 1] += h * ax; x[:, 3] += h * ay
    return x


def lmpc_config4(B=16384, seed=3):
    """BASELINE config 4 initial conditions: state, target, initial 34-parameter vectors."""
    rng = np.random.default_rng(seed)
    state = np.zeros((B, 8))
    state[:, 0] = rng.uniform(-0.1, 0.1, B)
    state[:, 2] = rng.uniform(-0.1, 0.1, B)
    state[:, 1] = rng.uniform(-0.1, 0.1, B)
    state[:, 3] = rng.uniform(-0.1, 0.1, B)
    target = np.zeros((B, 8))
    target[:, 0] = rng.uniform(-0.1, 0.1, B)
    target[:, 2] = rng.uniform(-0.1, 0.1, B)
    pvec = np.clip(1.0 + 0.1 * rng.standard_normal((B, 34)), 0.01, 1.9)
    u_prev = np.zeros((B, 2))
    return dict(state=state, target=target, pvec=pvec, u_prev=u_prev)


# ----------------------------------------------------------------------------- solver-ready input sets
def _governor(r_v, target, dr_max=0.01, alpha_rg=0.5):
    """rob_ctrl.py:346-348."""
    r_v = r_v.copy()
    for i in (0, 2):
        r_v[:, i] = r_v[:, i] + alpha_rg * np.clip(target[:, i] - r_v[:, i], -dr_max, dr_max)
    return r_v


def _ref_traj(r_v, target, N=20, step_fraction=0.2):
    """np_mpc_adaptive_with_linear_regressor.py:201-210, batched."""
    B = r_v.shape[0]
    R = np.zeros((B, N + 1, 4))
    for i in range(N + 1):
        w = 1.0 - (1.0 - step_fraction) ** (i + 1)
        r_i = r_v + w * (target - r_v)
        R[:, i, 0] = r_i[:, 0]
        R[:, i, 2] = r_i[:, 2]
    return R.reshape(B, -1)


def pmpc_inputs(states_per_object=4, seed=1):
    """(config dict, aux [B,4] = per-instance Qp, Qv, R, mu) for dart_solve."""
    c = pmpc_config2(states_per_object, seed)
    return c, np.ascontiguousarray(np.stack([c["Qp"], c["Qv"], c["R"], c["mu"]], axis=1))


def rmpc_inputs(B=32, seed=2, cold=False):
    """Mid-episode RMPC solver inputs: friction-like theta_hat, nonzero u_prev, governor-built staged reference."""
    c = rmpc_config3(B, seed)
    rng = np.random.default_rng(seed + 100)
    th = 0.01 * rng.standard_normal((B, 14))
    if not cold:
        th[:, 1] -= rng.uniform(0, 0.5, B); th[:, 4] -= rng.uniform(0, 1.0, B)
        th[:, 10] -= rng.uniform(0, 0.5, B); th[:, 12] -= rng.uniform(0, 1.0, B)
    up = np.zeros((B, 2)) if cold else rng.uniform(-0.3, 0.3, (B, 2))
    rv = np.zeros((B, 4)); rv[:, [0, 2]] = c["x0"][:, [0, 2]]
    rv = _governor(rv, c["target"])
    ref = _ref_traj(rv, c["target"])
    return dict(x0=c["x0"], ref=ref, aux=np.concatenate([up, th], axis=1), u_prev=up, theta=th)


def lmpc_inputs(B=32, seed=3):
    c = lmpc_config4(B, seed)
    rng = np.random.default_rng(seed + 100)
    up = rng.uniform(-0.2, 0.2, (B, 2))
    return dict(x0=c["state"], ref=c["target"], aux=np.concatenate([up, c["pvec"]], axis=1), u_prev=up, pvec=c["pvec"])


def arm_dynamics(B, seed=0, stress=1.0):
    """Synthetic but physically shaped inputs of the low-level arm QP (the reference gets them from MuJoCo, arm.py:111-200): SPD mass matrix, full-rank 6x7
    Jacobian, Mx_inv = J M^-1 J' as in arm.py:139, small pose errors.  ``stress`` scales velocities / errors so that
    torque and velocity rows become active."""
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((B, 7, 7))
    M = np.einsum('bij,bkj->bik', A, A) * 0.15 + np.eye(7) * np.array([2.0, 2.0, 1.0, 1.0, 0.5, 0.3, 0.2])
    jac = rng.standard_normal((B, 6, 7)) * 0.4
    jacDot = rng.standard_normal((B, 6, 7)) * 0.2 * stress
    Minv = np.linalg.inv(M)
    Mx_inv = jac @ Minv @ np.transpose(jac, (0, 2, 1))
    q = rng.uniform(-1.0, 1.0, (B, 7)); q[:, 3] = rng.uniform(0.2, 2.5, B)
    ee = rng.uniform(-0.3, 0.3, (B, 3))
    tau_lim = np.array([50.0, 50, 30, 30, 30, 20, 20])
    return dict(q=q, qd=rng.standard_normal((B, 7)) * 0.1 * stress, qdd_prev=rng.standard_normal((B, 7)) * 2.0,
                mocap_pos=ee + 2e-4 * stress * rng.standard_normal((B, 3)), ee_pos=ee,
                rotvec=rng.standard_normal((B, 3)) * 2e-3 * stress, jac=jac, jacDot=jacDot, M=M,
                h=rng.uniform(-0.5, 0.5, (B, 7)) * tau_lim * min(1.0, stress), Mx_inv=Mx_inv)
