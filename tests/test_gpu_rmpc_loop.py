"""RMPC online adaptation and closed loop on the GPU vs the oracle (RLS trajectory bar: 1e-6 relative)."""
import numpy as np
import pytest

import dart_b200
from oracle import ipm, problems, rls
from tests import helpers

pytestmark = pytest.mark.gpu


def test_rls_kernel_trajectory(built):
    import torch
    rng = np.random.default_rng(7)
    B, T = 64, 300
    dev = torch.device("cuda", 0)
    theta = torch.zeros((B, 2, 7), dtype=torch.float64, device=dev)
    P = (torch.eye(7, dtype=torch.float64, device=dev) * 1e3).repeat(B, 2, 1, 1).contiguous()
    th_o = np.zeros((B, 2, 7)); P_o = np.tile(np.eye(7) * 1e3, (B, 2, 1, 1))
    true = rng.standard_normal((B, 2, 7))
    worst = 0.0
    for t in range(T):
        x = 0.1 * rng.standard_normal((B, 4))
        phi = rls.regressor(x, 0.1)
        y = np.einsum("bep,bp->be", true, phi) + 1e-3 * rng.standard_normal((B, 2))
        dart_b200.rls_update_device(theta, P, torch.from_numpy(phi).to(dev), torch.from_numpy(y).to(dev), 0.995)
        th_o, P_o = rls.rls_update_batch(th_o, P_o, phi, y, 0.995)
        if t % 50 == 49 or t == T - 1:
            d = np.abs(theta.cpu().numpy() - th_o).max() / np.abs(th_o).max()
            worst = max(worst, d)
    assert worst <= helpers.TOL_RLS, worst
    assert np.abs(P.cpu().numpy() - P_o).max() <= 1e-6 * np.abs(P_o).max()


def test_rls_class_matches_reference_class(built):
    rng = np.random.default_rng(3)
    a, b = dart_b200.RLS(p=7, theta0=np.zeros(7), P0=1e3, lam=0.995), rls.RLS(7, np.zeros(7), 1e3, 0.995)
    for _ in range(40):
        phi = rng.standard_normal(7); y = rng.standard_normal()
        a.update(phi, y); b.update(phi, y)
    assert np.abs(a.get() - b.get()).max() <= 1e-9 * max(1.0, np.abs(b.get()).max())
    assert np.abs(a.P - b.P).max() <= 1e-9 * np.abs(b.P).max()


def _plant_step(x, u, mu_p, c_p, Ts=0.002, gz=-9.81):
    """Surrogate plant of SURVEY 8(d) config 3: v' = gz sin u - mu g tanh(v/.01) - c v, explicit Euler sub-steps."""
    x = x.copy()
    for _ in range(4):
        h = Ts / 4
        ax = gz * np.sin(u[:, 0]) - mu_p * 9.81 * np.tanh(x[:, 1] / 0.01) - c_p * x[:, 1]
        ay = gz * np.sin(u[:, 1]) - mu_p * 9.81 * np.tanh(x[:, 3] / 0.01) - c_p * x[:, 3]
        x[:, 0] += h * x[:, 1]; x[:, 2] += h * x[:, 3]
        x[:, 1] += h * ax; x[:, 3] += h * ay
    return x


def test_closed_loop_matches_oracle_loop(built):
    """T closed-loop steps of rob_ctrl.py's body: device-resident RMPCBatch vs the oracle doing the same on the CPU."""
    import torch
    B, T = 16, 12
    c = dart_b200.workloads.rmpc_config3(B, seed=2)
    x = c["x0"].copy(); x[:, [1, 3]] *= 0.5
    dev = torch.device("cuda", 0)
    # P0 = 1 keeps the early RLS transients small enough that every NLP of the episode stays feasible (with the
    # reference's P0 = 1e3 the first few estimates make the velocity-capped NLP infeasible: see the next test)
    P0 = 1.0
    ctl = dart_b200.RMPCBatch(B, c["target"], x, device=0, rls_P0=P0)
    rv0 = np.zeros((B, 4)); rv0[:, [0, 2]] = x[:, [0, 2]]
    ctl.set_virtual_reference(rv0)
    # oracle-side state
    th = np.zeros((B, 2, 7)); P = np.tile(np.eye(7) * P0, (B, 2, 1, 1))
    r_v = rv0.copy(); prev = x.copy(); u_prev = np.zeros((B, 2)); Xw = None; Uw = None
    x_o = x.copy()
    for t in range(T):
        u_gpu = ctl.step(torch.from_numpy(x).to(dev)).cpu().numpy()
        # oracle: same step on the oracle's own state trajectory
        phi = rls.regressor(prev, 0.1); y = rls.accel_measurement(x_o, prev, 0.002)
        th, P = rls.rls_update_batch(th, P, phi, y, 0.995)
        r_v = problems.reference_governor(r_v, c["target"])
        ref = problems.build_ref_traj(x_o, r_v, c["target"], 20, 4, 0.2)
        prob = problems.rmpc_problem(x_o, u_prev, th.reshape(B, 14), ref)
        if Xw is None:
            X0 = np.zeros((B, 21, 6)); U0 = np.zeros((B, 20, 2))        # reference: w0 = zeros at the first call
        else:
            X0, U0 = Xw, Uw
        # warm-started solves start the barrier at RMPCBatch's warm_mu (1e-4), the first one at 0.1
        sol = ipm.solve(prob, X0=X0, U0=U0, opts=ipm.Options(mu0=0.1 if t == 0 else ctl.warm_mu))
        assert (sol["status"] == 0).all()
        Xw, Uw = sol["X"], sol["U"]
        u_o = sol["U"][:, 0]
        assert np.abs(u_gpu - u_o).max() <= helpers.TOL_U0, (t, np.abs(u_gpu - u_o).max())
        dth = np.abs(ctl.theta.cpu().numpy() - th).max() / max(1e-12, np.abs(th).max())
        assert dth <= helpers.TOL_RLS, (t, dth)
        prev = x_o.copy(); u_prev = u_o.copy()
        x_o = _plant_step(x_o, u_o, c["mu_plant"], c["c_plant"])
        x = _plant_step(x, u_gpu, c["mu_plant"], c["c_plant"])
    assert (ctl.status.cpu().numpy() == 0).all()
    assert np.abs(x - x_o).max() < 1e-6


def test_infeasible_nlp_is_flagged_like_the_oracle(built):
    """Reference RLS start (P0 = 1e3): after two steps theta_hat ~ 20 makes the velocity caps unattainable.  The
    reference silently returns IPOPT's restoration iterate; here both solvers stop early and flag status 2."""
    import torch
    B = 16
    c = dart_b200.workloads.rmpc_config3(B, seed=2)
    x = c["x0"].copy(); x[:, [1, 3]] *= 0.5
    dev = torch.device("cuda", 0)
    ctl = dart_b200.RMPCBatch(B, c["target"], x, device=0)
    rv0 = np.zeros((B, 4)); rv0[:, [0, 2]] = x[:, [0, 2]]
    ctl.set_virtual_reference(rv0)
    for t in range(3):
        u = ctl.step(torch.from_numpy(x).to(dev)).cpu().numpy()
        if t == 2:
            theta = ctl.theta.cpu().numpy().reshape(B, 14)
            ref = ctl.ref.cpu().numpy(); aux = ctl.aux.cpu().numpy()
            sol = ipm.solve(problems.rmpc_problem(x, aux[:, :2], theta, ref), X0=None, U0=None)
            st = ctl.status.cpu().numpy()
            assert (st == 2).sum() >= 3 and ((st == 2) == (sol["status"] == 2)).all()
            assert ctl.iters.cpu().numpy().max() < 80          # early exit, not the 200-iteration cap
            ok = st == 0
            assert np.abs(u - sol["U"][:, 0])[ok].max() <= 1e-3   # different warm starts, same optimum where feasible
        x = _plant_step(x, u, c["mu_plant"], c["c_plant"])


def test_adaptive_class_dropin(built):
    model = dart_b200.GravityModel(-9.81, 0.002); data = dart_b200.StateHolder()
    ctl = dart_b200.AdaptiveNPMPCSmooth(model, data, Ts=0.002, nx=4, nu=2, N=20, Qp=80.0, Qv=2.0, Ru=0.02, Rdu=1.0,
                                        u_bounds=(-0.6, 0.6), du_bounds=(-0.06, 0.06), vmax=0.2, v_eps=0.1, target_body="object")
    assert ctl.gz == -9.81 and ctl.N == 20 and ctl.nx == 4
    rv = np.array([.005, 0, .005, 0]); tg = np.array([.05, 0, .05, 0])
    ref = ctl.build_ref_traj(np.zeros(4), rv, tg, ctl.N, ctl.nx, step_fraction=0.2)
    assert np.array_equal(ref, problems.build_ref_traj(None, rv[None], tg[None], 20, 4, 0.2)[0])
    u0, loss = ctl.solve(np.zeros(4), np.zeros(2), np.zeros(14), ref)
    assert loss.shape == (1,) and abs(loss[0] - 6.087598328) < 1e-7 and np.allclose(u0, [-0.0502635, -0.0502635], atol=1e-6)
    assert ctl.w0.shape == (124,) and np.array_equal(ctl.w0[84:86], u0)
