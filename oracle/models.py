"""Literal float64 restatements of the three reference plant models.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  Every function accepts
arrays with arbitrary leading batch dimensions and works for complex inputs so
that ``oracle.ipm`` can differentiate it by the complex-step method; nothing in
here contains a hand-derived derivative.
"""
import numpy as np


def _cabs(z):
    """|z| that stays holomorphic off the kink (CasADi ``fabs``; d|z|/dz = sign(z), sign(0)=0)."""
    return z * np.sign(np.real(z))


# ----------------------------------------------------------------------------- PMPC
def pmpc_dynamics(x, u, g, mu, Ts):
    """``PMPC._dynamics`` (PMPC/src/controller/mpc_3d.py:87-97).

    x = [px, vx, py, vy, pz, vz], u = [theta_x, theta_y]; ``g`` is
    ``model.opt.gravity[2]`` (negative, mpc_3d.py:23).
    """
    vx, vy, vz = x[..., 1], x[..., 3], x[..., 5]
    tx, ty = u[..., 0], u[..., 1]
    ax = g * np.sin(tx) - mu * vx
    ay = g * np.sin(ty) - mu * vy
    vz_new = -g * (tx ** 2 + ty ** 2)
    az = (vz_new - vz) / Ts
    return np.stack([vx, ax, vy, ay, vz_new, az], axis=-1)


def pmpc_step(x, u, g, mu, Ts):
    """``PMPC._rk4_step`` (mpc_3d.py:99-104)."""
    k1 = pmpc_dynamics(x, u, g, mu, Ts)
    k2 = pmpc_dynamics(x + Ts / 2 * k1, u, g, mu, Ts)
    k3 = pmpc_dynamics(x + Ts / 2 * k2, u, g, mu, Ts)
    k4 = pmpc_dynamics(x + Ts * k3, u, g, mu, Ts)
    return x + Ts / 6 * (k1 + 2 * k2 + 2 * k3 + k4)


# ----------------------------------------------------------------------------- RMPC
def rmpc_phi(x, v_eps):
    """``AdaptiveNPMPCSmooth._phi`` (np_mpc_adaptive_with_linear_regressor.py:171-176)."""
    px, vx, py, vy = x[..., 0], x[..., 1], x[..., 2], x[..., 3]
    one = np.ones_like(px)
    return np.stack([px, vx, py, vy, np.tanh(vx / v_eps), np.tanh(vy / v_eps), one], axis=-1)


def rmpc_dynamics(x, u, th, gz, v_eps):
    """``_dyn_regressor`` (np_mpc_adaptive_with_linear_regressor.py:178-186). th = [theta_x(7), theta_y(7)]."""
    vx, vy = x[..., 1], x[..., 3]
    phi = rmpc_phi(x, v_eps)
    ax = gz * np.sin(u[..., 0]) + np.sum(phi * th[..., 0:7], axis=-1)
    ay = gz * np.sin(u[..., 1]) + np.sum(phi * th[..., 7:14], axis=-1)
    return np.stack([vx, ax, vy, ay], axis=-1)


def rmpc_step(x, u, th, gz, v_eps, Ts):
    """``_rk4_step_regressor`` (np_mpc_adaptive_with_linear_regressor.py:188-193)."""
    k1 = rmpc_dynamics(x, u, th, gz, v_eps)
    k2 = rmpc_dynamics(x + Ts / 2 * k1, u, th, gz, v_eps)
    k3 = rmpc_dynamics(x + Ts / 2 * k2, u, th, gz, v_eps)
    k4 = rmpc_dynamics(x + Ts * k3, u, th, gz, v_eps)
    return x + Ts / 6 * (k1 + 2 * k2 + 2 * k3 + k4)


# ----------------------------------------------------------------------------- LMPC
def _squash(p):
    """``squash_param`` (LMPC/src/controller/rlmpc2.py:287-289): |p| + 1e-6."""
    return _cabs(p) + 1e-6


def _stribeck(v, F_s, F_c, B, v_s, eps):
    """``stribeck_fric`` (rlmpc2.py:355-359)."""
    abs_v = _cabs(v)
    exp_term = np.exp(-abs_v / (v_s + 1e-12))
    return np.tanh(v / eps) * (F_c + (F_s - F_c) * exp_term) + B * v


def lmpc_dynamics(x, u, pvec, g=9.81):
    """``safe_dynamics`` (rlmpc2.py:260-429); the index map is the code's, not the docstring's.

    x = [px, vx, py, vy, theta_x, omega_x, theta_y, omega_y], u = [a, b], pvec (34).
    ``g`` is the literal 9.81 of rlmpc2.py:342 (``packet["g"]`` is never read).
    """
    P = lambda i: pvec[..., i]
    px, vx, py, vy = x[..., 0], x[..., 1], x[..., 2], x[..., 3]
    theta_x, omega_x, theta_y, omega_y = x[..., 4], x[..., 5], x[..., 6], x[..., 7]
    a, b = u[..., 0], u[..., 1]

    m_x, m_y = _squash(P(0)), _squash(P(1))
    c_x, c_y = _squash(P(2)), _squash(P(3))
    k_x, k_y = _squash(P(4)), _squash(P(5))
    F_s_x, F_c_x, B_x = P(6), P(7), P(8)
    v_s_x, eps_x = _squash(P(9)), _squash(P(10))
    F_s_y, F_c_y, B_y = P(11), P(12), P(13)
    v_s_y, eps_y = _squash(P(14)), _squash(P(15))
    I_x, I_y = _squash(P(16)), _squash(P(17))
    r_x, r_y = _squash(P(18)), _squash(P(19))
    c_rot_x, c_rot_y = _squash(P(20)), _squash(P(21))
    F_s_rot_x, F_c_rot_x, B_rot_x = P(22), P(23), P(24)
    v_s_rot_x, eps_rot_x = _squash(P(25)), _squash(P(26))
    F_s_rot_y, F_c_rot_y, B_rot_y = P(27), P(28), P(29)
    v_s_rot_y, eps_rot_y = _squash(P(30)), _squash(P(31))
    h_com_x, h_com_y = _squash(P(32)), _squash(P(33))

    G_x = m_x * (g * np.sin(a))
    G_y = m_y * (g * np.sin(b))

    Ff_x = _stribeck(vx, F_s_x, F_c_x, B_x, v_s_x, eps_x)
    Ff_y = _stribeck(vy, F_s_y, F_c_y, B_y, v_s_y, eps_y)

    v_slip_x = vx - r_x * omega_y
    v_slip_y = vy - (-r_y * omega_x)
    F_roll_x = _stribeck(v_slip_x, F_s_x, F_c_x, B_x, v_s_x, eps_x)
    F_roll_y = _stribeck(v_slip_y, F_s_y, F_c_y, B_y, v_s_y, eps_y)

    tau_slip_x = -r_y * F_roll_y
    tau_slip_y = -r_x * F_roll_x
    T_noslip_x = _stribeck(omega_x, F_s_rot_x, F_c_rot_x, B_rot_x, v_s_rot_x, eps_rot_x)
    T_noslip_y = _stribeck(omega_y, F_s_rot_y, F_c_rot_y, B_rot_y, v_s_rot_y, eps_rot_y)
    T_damp_x = c_rot_x * omega_x
    T_damp_y = c_rot_y * omega_y
    tau_topple_x = -m_y * g * h_com_x * np.sin(theta_x)
    tau_topple_y = -m_x * g * h_com_y * np.sin(theta_y)
    tau_x = tau_slip_x - T_noslip_x - T_damp_x + tau_topple_x
    tau_y = tau_slip_y - T_noslip_y - T_damp_y + tau_topple_y
    alpha_rot_x = tau_x / (I_x + 1e-12)
    alpha_rot_y = tau_y / (I_y + 1e-12)

    rhs_x = G_x - c_x * vx - k_x * px - Ff_x - F_roll_x
    rhs_y = G_y - c_y * vy - k_y * py - Ff_y - F_roll_y
    qdd_x = rhs_x / m_x
    qdd_y = rhs_y / m_y
    return np.stack([vx, qdd_x, vy, qdd_y, omega_x, alpha_rot_x, omega_y, alpha_rot_y], axis=-1)


def lmpc_step(x, u, pvec, Ts):
    """``_rk4`` (rlmpc2.py:431-436)."""
    k1 = lmpc_dynamics(x, u, pvec)
    k2 = lmpc_dynamics(x + 0.5 * Ts * k1, u, pvec)
    k3 = lmpc_dynamics(x + 0.5 * Ts * k2, u, pvec)
    k4 = lmpc_dynamics(x + Ts * k3, u, pvec)
    return x + Ts * (k1 + 2 * k2 + 2 * k3 + k4) / 6


# ----------------------------------------------------------------------------- tilt -> quaternion
def tilt_to_quat(u):
    """Euler xyz [u1, -u0, 0] -> quaternion wxyz (PMPC/main.py:107-116; rob_ctrl.py:355; run.py:259-261)."""
    u = np.asarray(u, dtype=np.float64)
    ang = np.stack([u[..., 1], -u[..., 0], np.zeros_like(u[..., 0])], axis=-1)
    c = np.cos(ang / 2.0)
    s = np.sin(ang / 2.0)
    cx, cy, cz = c[..., 0], c[..., 1], c[..., 2]
    sx, sy, sz = s[..., 0], s[..., 1], s[..., 2]
    return np.stack([cx * cy * cz + sx * sy * sz,
                     sx * cy * cz - cx * sy * sz,
                     cx * sy * cz + sx * cy * sz,
                     cx * cy * sz - sx * sy * cz], axis=-1)
