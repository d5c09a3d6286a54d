"""Low-level arm QP on the GPU against the CPU oracle.  The QP is strictly convex, so the minimiser is
unique; the bar is |dx| <= 1e-6 (1 + |x|) on the accelerations (tol 1e-8 on both sides) and 1e-9 relative on the cost."""
import numpy as np
import pytest

import dart_b200
from dart_b200 import arm as parm
from oracle import arm_qp

pytestmark = pytest.mark.gpu


def _solve_gpu(H, g, C, lo, hi, x0=None, **kw):
    import torch
    t = lambda a: None if a is None else torch.from_numpy(np.ascontiguousarray(a)).cuda()
    out = parm.solve_qp_device(t(H), t(g), t(C), t(lo), t(hi), x0=t(x0), **kw)
    torch.cuda.synchronize()
    return {k: v.cpu().numpy() for k, v in out.items()}


@pytest.mark.parametrize("stress,B", [(0.3, 257), (1.0, 1000), (3.0, 1000), (10.0, 300)])
def test_parity_with_oracle(built, stress, B):
    dyn = dart_b200.workloads.arm_dynamics(B, seed=11, stress=stress)
    H, g, c0, C, lo, hi = arm_qp.build_qp(dyn, arm_qp.default_params())
    ref = arm_qp.solve_qp(H, g, C, lo, hi)
    out = _solve_gpu(H, g, C, lo, hi)
    assert (ref["status"] == 0).all() and (out["status"] == 0).all()
    dx = np.abs(out["x"] - ref["x"]).max(axis=1) / (1.0 + np.abs(ref["x"]).max(axis=1))
    dj = np.abs(out["obj"] - ref["obj"]) / (1.0 + np.abs(ref["obj"]))
    print(f"stress {stress}: max rel |dx| {dx.max():.2e}, rel dJ {dj.max():.2e}, iters gpu {out['iters'].mean():.1f} oracle {ref['iters'].mean():.1f}")
    assert dx.max() < 1e-6 and dj.max() < 1e-9
    assert np.abs(out["iters"] - ref["iters"]).max() <= 1                      # same method: iteration for iteration
    stat, viol = arm_qp.kkt_residual(H[:64], g[:64], C[:64], lo[:64], hi[:64], out["x"][:64])
    gn = np.linalg.norm(np.einsum('bij,bj->bi', H[:64], out["x"][:64]) + g[:64], axis=1) + 1.0
    assert (stat / gn).max() < 1e-6 and viol.max() < 1e-7


def test_edge_cases(built):
    dyn = dart_b200.workloads.arm_dynamics(9, seed=5, stress=1.0)
    H, g, c0, C, lo, hi = arm_qp.build_qp(dyn, arm_qp.default_params())
    hi[4, 16] = lo[4, 16] - 1.0                                                # empty torque interval
    out = _solve_gpu(H, g, C, lo, hi)
    assert out["status"][4] == dart_b200.STATUS_INFEASIBLE and (np.delete(out["status"], 4) == 0).all()
    capped = _solve_gpu(H, g, C, lo, hi, max_iter=2)
    assert (np.delete(capped["status"], 4) == dart_b200.STATUS_MAXITER).all() and (np.delete(capped["iters"], 4) == 2).all()
    # B = 1 and B not a multiple of the 16 QPs of a block
    one = _solve_gpu(H[:1], g[:1], C[:1], lo[:1], hi[:1])
    assert np.array_equal(one["x"][0], out["x"][0])
    # primal warm start from the solution: same optimum
    warm = _solve_gpu(H, g, C, lo, hi, x0=np.where(out["status"][:, None] == 0, out["x"], 0.0))
    ok = out["status"] == 0
    assert np.abs(warm["x"][ok] - out["x"][ok]).max() < 1e-6 * (1 + np.abs(out["x"][ok]).max())


def test_batch_wrapper_returns_torque_and_reference_loss(built):
    P = parm.default_params()
    dyn = dart_b200.workloads.arm_dynamics(64, seed=3, stress=1.0)
    ctl = dart_b200.ArmQPBatch(P, device=0)
    tau, loss, x = ctl.solve(dyn)
    H, g, c0, C, lo, hi = arm_qp.build_qp(dyn, arm_qp.default_params())
    ref = arm_qp.solve_qp(H, g, C, lo, hi)
    assert np.abs(x - ref["x"]).max() < 1e-6 * (1 + np.abs(ref["x"]).max())
    assert np.abs(loss - (ref["obj"] + c0)).max() <= 1e-8 * (1 + np.abs(ref["obj"] + c0).max())
    tau_ref = np.einsum('bij,bj->bi', dyn["M"], ref["x"]) + dyn["h"]
    assert np.abs(tau - tau_ref).max() < 1e-5
    assert (tau <= P["taumax"] + 1e-6).all() and (tau >= P["taumin"] - 1e-6).all()      # torque limits hold (arm.py:401-405)
    # second cycle: warm start from the previous accelerations
    tau2, loss2, x2 = ctl.solve(dyn)
    assert np.abs(x2 - x).max() < 1e-6 * (1 + np.abs(x).max()) and ctl.launches == 4      # QP build + QP solve per cycle


def test_facade_single_arm(built):
    P = parm.default_params()
    P["joint_names"] = [f"R_joint{i}" for i in range(1, 8)]
    dyn = dart_b200.workloads.arm_dynamics(1, seed=8, stress=0.5)
    arm = dart_b200.ARMCONTROL(None, None, P)
    tau, loss = arm.compute_torque_from({k: v[0] for k, v in dyn.items()})
    H, g, c0, C, lo, hi = arm_qp.build_qp(dyn, arm_qp.default_params())
    ref = arm_qp.solve_qp(H, g, C, lo, hi)
    assert np.abs(tau - (dyn["M"][0] @ ref["x"][0] + dyn["h"][0])).max() < 1e-5
    assert abs(loss - (ref["obj"][0] + c0[0])) <= 1e-8 * (1 + abs(loss))


@pytest.mark.parametrize("stress", [0.3, 3.0])
def test_device_qp_build_matches_literal_restatement(built, stress):
    """dart_arm_qp_build (Jacobi eigen-decompositions on the GPU) against the per-instance numpy restatement of arm.py:337-405."""
    import torch
    P = parm.default_params()
    P["Wsmooth"] = np.eye(7) * 1e-7
    dyn = dart_b200.workloads.arm_dynamics(200, seed=21, stress=stress)
    dyn["Mx_inv"][3] *= 1e-3                      # |det| < 1e-8: the pinv(rcond=1e-3) branch of arm.py:351-357
    dyn["Mx_inv"][5] = dyn["Mx_inv"][5] @ np.diag([1, 1, 1, 1, 1, 1e-9]) ; dyn["Mx_inv"][5] = 0.5 * (dyn["Mx_inv"][5] + dyn["Mx_inv"][5].T)
    Pref = arm_qp.default_params(); Pref["Wsmooth"] = P["Wsmooth"]
    ref = arm_qp.build_qp(dyn, Pref)
    d = {k: torch.from_numpy(np.ascontiguousarray(v)).cuda() for k, v in dyn.items()}
    out = parm.build_qp_device(d, P)
    torch.cuda.synchronize()
    for name, b in zip(("H", "g", "c0", "C", "lo", "hi"), ref):
        a = out[name].cpu().numpy()
        scale = np.abs(b).reshape(b.shape[0], -1).max(axis=1) + 1.0
        err = (np.abs(a - b).reshape(b.shape[0], -1).max(axis=1) / scale).max()
        print(f"{name}: max rel err {err:.2e}")
        assert err < 1e-9, name
