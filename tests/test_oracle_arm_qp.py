"""CPU checks of the arm-QP oracle: the interior-point solution against scipy's SLSQP and against the KKT
conditions themselves, and the product's vectorised QP build against the literal, per-instance restatement."""
import numpy as np
import pytest
from scipy.optimize import minimize

import dart_b200
from dart_b200 import arm as parm
from oracle import arm_qp


@pytest.mark.parametrize("stress", [0.3, 1.0, 3.0])
def test_ipm_matches_slsqp_and_kkt(stress):
    P = arm_qp.default_params()
    dyn = dart_b200.workloads.arm_dynamics(12, seed=4, stress=stress)
    H, g, c0, C, lo, hi = arm_qp.build_qp(dyn, P)
    out = arm_qp.solve_qp(H, g, C, lo, hi)
    assert (out["status"] == 0).all()
    stat, viol = arm_qp.kkt_residual(H, g, C, lo, hi, out["x"])
    gnorm = np.linalg.norm(np.einsum('bij,bj->bi', H, out["x"]) + g, axis=1) + 1.0
    assert (stat / gnorm).max() < 1e-6 and viol.max() < 1e-9
    for b in range(4):
        cons = [dict(type="ineq", fun=lambda x, b=b: C[b] @ x - lo[b], jac=lambda x, b=b: C[b]),
                dict(type="ineq", fun=lambda x, b=b: hi[b] - C[b] @ x, jac=lambda x, b=b: -C[b])]
        r = minimize(lambda x: 0.5 * x @ H[b] @ x + g[b] @ x, np.zeros(7), jac=lambda x: H[b] @ x + g[b], method="SLSQP",
                     constraints=cons, options=dict(ftol=1e-13, maxiter=500))
        assert r.status in (0, 8)      # 8 = SLSQP's line search found no further decrease at its tolerance: also an optimum
        # strictly convex: unique minimiser
        assert abs(r.fun - out["obj"][b]) <= 1e-7 * (1.0 + abs(r.fun))
        assert np.abs(r.x - out["x"][b]).max() < 1e-4 * (1.0 + np.abs(r.x).max())


def test_vectorised_build_matches_literal_restatement():
    P = arm_qp.default_params()
    dyn = dart_b200.workloads.arm_dynamics(16, seed=9, stress=2.0)
    dyn["Mx_inv"][3] *= 1e-3                      # |det| < 1e-8: the pinv branch of arm.py:351-357
    ref = arm_qp.build_qp(dyn, P)
    out = parm.build_qp(dyn, parm.default_params())
    for a, b, name in zip(out, ref, ("H", "g", "c0", "C", "lo", "hi")):
        scale = np.abs(b).max() + 1.0
        assert np.abs(a - b).max() <= 1e-9 * scale, name


def test_loss_is_the_reference_cost():
    """0.5 x'Hx + g'x + c0 equals Eimp'Wimp Eimp + Epos'Wpos Epos + qddd'Wsmooth qddd (arm.py:392-396) at any x."""
    P = arm_qp.default_params()
    P["Wsmooth"] = np.eye(7) * 1e-6
    dyn = dart_b200.workloads.arm_dynamics(3, seed=2)
    H, g, c0, C, lo, hi = arm_qp.build_qp(dyn, P)
    rng = np.random.default_rng(0)
    for b in range(3):
        d = {k: v[b] for k, v in dyn.items()}
        x = rng.standard_normal(7) * 10
        Minv = np.linalg.pinv(d["M"], rcond=1e-6); Mx = np.linalg.inv(d["Mx_inv"])
        mu = Mx @ (d["jac"] @ (Minv @ d["h"]) + d["jacDot"] @ d["qd"])
        D = arm_qp.safe_matrix_sqrt(Mx) @ np.sqrt(P["K"]) + np.sqrt(P["K"]) @ arm_qp.safe_matrix_sqrt(Mx)
        twist = np.concatenate([d["mocap_pos"] - d["ee_pos"], d["rotvec"]])
        F = -D @ (d["jac"] @ d["qd"]) + P["K"] @ twist + mu
        Eimp = d["jac"] @ x + d["jacDot"] @ d["qd"] - d["Mx_inv"] @ F
        Epos = x - (2.0 * np.sqrt(np.diag(P["K_null"])) * (-d["qd"]) + P["K_null"] @ (-d["q"]))
        qddd = (x - d["qdd_prev"]) / P["dt"]
        cost = Eimp @ P["Wimp"] @ Eimp + Epos @ P["Wpos"] @ Epos + qddd @ P["Wsmooth"] @ qddd
        mine = 0.5 * x @ H[b] @ x + g[b] @ x + c0[b]
        assert abs(cost - mine) <= 1e-9 * (1.0 + abs(cost))


def test_empty_box_flagged_infeasible():
    P = arm_qp.default_params()
    dyn = dart_b200.workloads.arm_dynamics(4, seed=1)
    H, g, c0, C, lo, hi = arm_qp.build_qp(dyn, P)
    hi[2, 5] = lo[2, 5] - 1.0
    out = arm_qp.solve_qp(H, g, C, lo, hi)
    assert out["status"][2] == arm_qp.STATUS_INFEASIBLE and (np.delete(out["status"], 2) == 0).all()
