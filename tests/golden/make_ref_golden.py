"""Reference-generated fixtures: run the reference's OWN source (``/root/reference``) and record what it computes.

Run here (the container that has /root/reference):   python tests/golden/make_ref_golden.py
Writes tests/golden/ref_pmpc.npz, ref_rmpc.npz, ref_lmpc.npz, ref_policy.npz and tests/golden/tapes/*.npz.

How the reference runs without CasADi / MuJoCo: ``oracle/refshim`` (a graph-recording ``casadi`` stand-in, duck-typed
MuJoCo objects, stepping gates for the worker loops).  Every number below is produced by executing the reference's
unmodified files: ``PMPC/src/controller/mpc_3d.py``, ``RMPC/dev_dual/controller/np_mpc_adaptive_with_linear_regressor.py``,
``LMPC/src/controller/rlmpc2.py``.  The NLP *solutions* come from ``oracle/refshim/nlp_solve.py`` applied to the
reference's captured problems at 1e-10 (NOT from IPOPT, which cannot be installed; the KKT point of these NLPs does not
depend on the solver -- the tests re-verify every stored solution's KKT residual on the reference's own graph).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle.refshim import casadi as shim, harness, loader          # noqa: E402
from oracle import models as omodels                                 # noqa: E402  (surrogate plant step of the facade trace only)
import dart_b200                                                     # noqa: E402  (workloads only: seeded inputs)

TAPES = os.path.join(HERE, "tapes")
W = dart_b200.workloads


def nlp_probe(tape, rng, n_pts, w_scale, p_fn):
    """Reference NLP functions at seeded random points (not solutions): f, g, grad f, Jacobian of g."""
    nw, np_ = len(tape.inputs["x"]), len(tape.inputs["p"])
    w = rng.standard_normal((n_pts, nw)) * w_scale
    p = np.stack([p_fn(rng) for _ in range(n_pts)])
    ev = tape.eval(x=w, p=p)
    gradf = tape.grad("f", "x", x=w, p=p)
    jg = tape.jac("g", "x", x=w, p=p)
    return dict(w=w, p=p, f=ev["f"][:, 0], g=ev["g"], gradf=gradf, jg=jg)


def solve_record(calls):
    keys = ("x", "f", "lam_g", "lam_x", "iters", "kkt", "status", "p", "x0")
    return {k: np.array([c[k] for c in calls]) for k in keys}


# ----------------------------------------------------------------------------------------------------------------------
def make_pmpc():
    rng = np.random.default_rng(11)
    out = {}
    # a1/a2: the reference's discrete dynamics Function (mpc_3d.py:30) for each friction value of config 2
    dx, du, dmu, dout = [], [], [], []
    for mu in W.PMPC_FRICTIONS:
        mpc = harness.pmpc(nx=6, nu=2, N=15, Qp=600.0, Qv=5.0, R=0.1, mu=mu, u_bounds=(-0.6, 0.6))
        for _ in range(16):
            x = rng.uniform(-0.3, 0.3, 6)
            x[4] = 0.43 + 0.05 * rng.standard_normal()
            u = rng.uniform(-0.7, 0.7, 2)
            dx.append(x); du.append(u); dmu.append(mu)
            dout.append(mpc.f(x, u).full().ravel())
    out.update(dyn_x=np.array(dx), dyn_u=np.array(du), dyn_mu=np.array(dmu), dyn_out=np.array(dout))

    # a3/a4: the reference's NLP per (shape weights, friction) and PMPC.solve on config-2 instances (2 per object)
    c = W.pmpc_config2(states_per_object=2, seed=1)
    B = len(c["mu"])
    combos = {}
    u0, J, w_opt, lam_g, iters, kkt, combo_id = [], [], [], [], [], [], []
    probes = {k: [] for k in ("w", "p", "f", "g", "gradf", "jg", "combo")}
    for i in range(B):
        key = (c["Qp"][i], c["Qv"][i], c["R"][i], c["mu"][i])
        if key not in combos:
            mpc = harness.pmpc(nx=6, nu=2, N=15, Qp=key[0], Qv=key[1], R=key[2], mu=key[3], u_bounds=(-0.6, 0.6))
            cid = len(combos)
            combos[key] = (cid, mpc)
            mpc.solver.tape.meta.update(Qp=key[0], Qv=key[1], R=key[2], mu=key[3], N=15,
                                        lbx=np.array(mpc.lbx), ubx=np.array(mpc.ubx), source="mpc_3d.py:28-85")
            mpc.solver.tape.save(os.path.join(TAPES, f"pmpc_nlp_{cid}.npz"))
            pr = nlp_probe(mpc.solver.tape, rng, 2, 0.2, lambda r: np.concatenate([r.uniform(-.2, .2, 6), r.uniform(-.2, .2, 6)]))
            for k in pr:
                probes[k].append(pr[k])
            probes["combo"].append(np.full(2, cid))
        cid, mpc = combos[key]
        harness.set_pmpc_state(mpc, c["state"][i])
        assert np.array_equal(mpc.get_state(), c["state"][i])          # a5
        u, loss = mpc.solve(c["target"][i])
        call = mpc.solver.calls[-1]
        assert call["status"] == 0, (i, call["kkt"])
        assert np.array_equal(mpc.w0, call["x"])
        u0.append(u); J.append(loss[0]); w_opt.append(call["x"]); lam_g.append(call["lam_g"])
        iters.append(call["iters"]); kkt.append(call["kkt"]); combo_id.append(cid)
        print(f"pmpc {i:3d} combo {cid} J={loss[0]:.9f} u0={u} it={call['iters']} kkt={call['kkt']:.1e}", flush=True)
    out.update(state=c["state"], target=c["target"], Qp=c["Qp"], Qv=c["Qv"], R=c["R"], mu=c["mu"],
               u0=np.array(u0), J=np.array(J), w=np.array(w_opt), lam_g=np.array(lam_g), iters=np.array(iters),
               kkt=np.array(kkt), combo=np.array(combo_id),
               combo_params=np.array([k for k in combos]))
    out.update({"probe_" + k: np.concatenate(v) for k, v in probes.items()})

    # config 1 (PMPC/main.py:59-69 weights) on the README example
    c1 = W.pmpc_config1()
    mpc = harness.pmpc(nx=6, nu=2, N=15, Qp=400.0, Qv=2.0, R=0.2, mu=0.10, u_bounds=(-0.6, 0.6))
    harness.set_pmpc_state(mpc, c1["state"][0])
    u, loss = mpc.solve(c1["target"][0])
    mpc.solver.tape.meta.update(Qp=400.0, Qv=2.0, R=0.2, mu=0.1, N=15, lbx=np.array(mpc.lbx), ubx=np.array(mpc.ubx),
                                source="mpc_3d.py:28-85 with PMPC/main.py:59-69")
    mpc.solver.tape.save(os.path.join(TAPES, "pmpc_nlp_config1.npz"))
    out.update(c1_u0=u, c1_J=loss, c1_w=mpc.solver.calls[-1]["x"], c1_lam_g=mpc.solver.calls[-1]["lam_g"])
    print("pmpc config1", u, loss)
    np.savez_compressed(os.path.join(HERE, "ref_pmpc.npz"), **out)


# ----------------------------------------------------------------------------------------------------------------------
RMPC_PARAMS = dict(nx=4, nu=2, N=20, Qp=80.0, Qv=2.0, Ru=0.02, Rdu=1.0, u_bounds=(-0.6, 0.6), du_bounds=(-0.06, 0.06),
                   vmax=0.2, v_eps=0.1, target_body="object")       # rob_ctrl.py:281-284


def make_rmpc():
    rng = np.random.default_rng(12)
    mod = loader.load(loader.RMPC_FILE)
    out = {}
    ctl = harness.rmpc(**RMPC_PARAMS)
    tape = ctl.solver.tape
    tape.meta.update(lbx=ctl.lbx, ubx=ctl.ubx, lbg=ctl.lbg, ubg=ctl.ubg, N=20,
                     source="np_mpc_adaptive_with_linear_regressor.py:65-168 with rob_ctrl.py:281-284")
    tape.save(os.path.join(TAPES, "rmpc_nlp.npz"))

    # a7: RLS class trajectories (reference class, np_mpc...:10-30), inputs shaped like rob_ctrl.py:335-343
    T, E = 96, 4
    phi = np.zeros((E, T, 7)); y = np.zeros((E, T)); th = np.zeros((E, T, 7)); Pm = np.zeros((E, T, 7, 7))
    for e in range(E):
        r = mod.RLS(p=7, theta0=np.zeros(7), P0=1e3, lam=0.995)
        st = rng.uniform(-0.1, 0.1, 4)
        th_true = np.array([0, -0.3, 0, 0, -0.8, 0, 0.02]) * (1 + e)
        for t in range(T):
            st = st + 0.02 * rng.standard_normal(4)
            ph = np.array([st[0], st[1], st[2], st[3], np.tanh(st[1] / 0.1), np.tanh(st[3] / 0.1), 1.0])
            yy = ph @ th_true + 0.05 * rng.standard_normal() if t else 0.0
            r.update(ph, yy)
            phi[e, t], y[e, t], th[e, t], Pm[e, t] = ph, yy, r.get(), r.P
    out.update(rls_phi=phi, rls_y=y, rls_theta=th, rls_P=Pm)

    # a10: build_ref_traj
    rv = rng.uniform(-0.1, 0.1, (8, 4)); tg = rng.uniform(-0.1, 0.1, (8, 4))
    out.update(ref_rv=rv, ref_target=tg,
               ref_out=np.array([ctl.build_ref_traj(None, rv[i], tg[i], 20, 4, step_fraction=0.2) for i in range(8)]))

    # a8: f_disc (np_mpc...:69-73)
    dx = rng.uniform(-0.3, 0.3, (32, 4)); du = rng.uniform(-0.7, 0.7, (32, 2)); dth = rng.standard_normal((32, 14)) * 0.5
    out.update(dyn_x=dx, dyn_u=du, dyn_th=dth,
               dyn_out=np.array([ctl.f_disc(dx[i], du[i], dth[i]).full().ravel() for i in range(32)]))

    # a9: NLP functions at random points
    pr = nlp_probe(tape, rng, 3, 0.2, lambda r: np.concatenate([r.uniform(-.1, .1, 4), r.uniform(-.3, .3, 2),
                                                                   r.standard_normal(14) * 0.3, r.uniform(-.1, .1, 84)]))
    out.update({"probe_" + k: v for k, v in pr.items()})

    # a11: AdaptiveNPMPCSmooth.solve on the mid-episode inputs the GPU tests use (fresh controller per instance => w0 = 0)
    d = W.rmpc_inputs(B=12, seed=2)
    u0, J = [], []
    calls = []
    for i in range(12):
        ci = harness.rmpc(**RMPC_PARAMS)
        u, loss = ci.solve(d["x0"][i], d["u_prev"][i], d["theta"][i], d["ref"][i])
        call = ci.solver.calls[-1]
        assert call["status"] == 0, (i, call["kkt"])
        calls.append(call); u0.append(u); J.append(loss[0])
        print(f"rmpc {i:3d} J={loss[0]:.9f} u0={u} it={call['iters']} kkt={call['kkt']:.1e}", flush=True)
    rec = solve_record(calls)
    out.update(x0=d["x0"][:12], u_prev=d["u_prev"][:12], theta=d["theta"][:12], ref=d["ref"][:12], u0=np.array(u0),
               J=np.array(J), w=rec["x"], lam_g=rec["lam_g"], iters=rec["iters"], kkt=rec["kkt"])

    # a7 + a10 + a11 + a12 together: the main loop of rob_ctrl.py:330-352 around the reference classes, on the
    # surrogate plant of config 3; warm start through ctl.w0 exactly as the reference carries it.
    c3 = W.rmpc_config3(B=3, seed=2)
    Ts, T = 0.002, 10
    gz = -9.81
    loop = {k: [] for k in ("x", "u0", "J", "theta", "r_v", "iters")}
    for i in range(3):
        ci = harness.rmpc(**RMPC_PARAMS)
        # P0 = 1 (reference: 1e3, rob_ctrl.py:286): with 1e3 the first estimates make the velocity-capped NLP infeasible
        # after two steps and the reference then returns IPOPT's restoration iterate, which no other solver reproduces
        rls_x = mod.RLS(p=7, theta0=np.zeros(7), P0=1.0, lam=0.995)
        rls_y = mod.RLS(p=7, theta0=np.zeros(7), P0=1.0, lam=0.995)
        theta_hat = np.zeros(14)
        xk = c3["x0"][i].copy()
        xk[[1, 3]] *= 0.5
        target = c3["target"][i].copy()
        r_v = np.array([xk[0], 0.0, xk[2], 0.0])        # the object starts at the tray centre in the reference; here r_v starts at the object
        u_prev = np.zeros(2)
        prev_state = xk.copy()
        dr_max, alpha_rg = 0.01, 0.5
        tr = {k: [] for k in loop}
        for t in range(T):
            ax_meas = (xk[1] - prev_state[1]) / Ts
            ay_meas = (xk[3] - prev_state[3]) / Ts
            phi_prev = np.array([prev_state[0], prev_state[1], prev_state[2], prev_state[3],
                                 np.tanh(prev_state[1] / ci.v_eps), np.tanh(prev_state[3] / ci.v_eps), 1.0])
            rls_x.update(phi_prev, ax_meas)
            rls_y.update(phi_prev, ay_meas)
            theta_hat[:7] = rls_x.get()
            theta_hat[7:] = rls_y.get()
            err_pos = np.array([target[0] - r_v[0], 0.0, target[2] - r_v[2], 0.0])
            step_pos = np.array([np.clip(err_pos[0], -dr_max, dr_max), 0.0, np.clip(err_pos[2], -dr_max, dr_max), 0.0])
            r_v = r_v + alpha_rg * step_pos
            Rref = ci.build_ref_traj(xk, r_v, target, ci.N, ci.nx, step_fraction=0.2)
            u_cmd, loss = ci.solve(xk, u_prev, theta_hat, Rref)
            call = ci.solver.calls[-1]
            tr["x"].append(xk.copy()); tr["u0"].append(u_cmd.copy()); tr["J"].append(loss[0])
            tr["theta"].append(theta_hat.copy()); tr["r_v"].append(r_v.copy()); tr["iters"].append(call["iters"])
            print(f"rmpc loop {i} t={t} J={loss[0]:.6f} u0={u_cmd} status={call['status']} kkt={call['kkt']:.1e}", flush=True)
            prev_state = xk.copy()
            u_prev = u_cmd.copy()
            xk = W.rmpc_plant_step(xk[None], u_cmd[None], c3["mu_plant"][i:i + 1], c3["c_plant"][i:i + 1], Ts, gz)[0]
        for k in loop:
            loop[k].append(np.array(tr[k]))
    out.update({"loop_" + k: np.array(v) for k, v in loop.items()})
    out.update(loop_x0=c3["x0"], loop_target=c3["target"], loop_mu_plant=c3["mu_plant"], loop_c_plant=c3["c_plant"])
    np.savez_compressed(os.path.join(HERE, "ref_rmpc.npz"), **out)


# ----------------------------------------------------------------------------------------------------------------------
def make_lmpc():
    rng = np.random.default_rng(13)
    out = {}
    # a13: safe_dynamics / _rk4 compiled from their own source text inside _solver_worker (rlmpc2.py:260-436)
    packet = dict(harness.LMPC_PACKET)
    safe_dynamics, span = loader.nested_function(loader.LMPC_FILE, ("RLMPC", "_solver_worker"), "safe_dynamics", {"ca": shim})
    rk4, span2 = loader.nested_function(loader.LMPC_FILE, ("RLMPC", "_solver_worker"), "_rk4",
                                        {"ca": shim, "safe_dynamics": safe_dynamics, "packet": packet})
    xs, us, ps = shim.SX.sym("x", 8), shim.SX.sym("u", 2), shim.SX.sym("pv", 34)
    f_c = shim.Function("safe_dynamics", [xs, us, ps], [safe_dynamics(xs, us, ps)])
    f_d = shim.Function("rk4", [xs, us, ps], [rk4(xs, us, ps)])
    n = 32
    dx = rng.uniform(-0.3, 0.3, (n, 8)); du = rng.uniform(-0.5, 0.5, (n, 2))
    dp = np.clip(1.0 + 0.4 * rng.standard_normal((n, 34)), 0.01, 1.9)
    dp[::4] *= np.where(rng.random((n // 4 + (n % 4 > 0), 34)) < 0.2, -1.0, 1.0)[: len(dp[::4])]      # some negative raw parameters (|p| squash)
    out.update(dyn_x=dx, dyn_u=du, dyn_p=dp,
               dyn_cont=np.array([f_c(dx[i], du[i], dp[i]).full().ravel() for i in range(n)]),
               dyn_out=np.array([f_d(dx[i], du[i], dp[i]).full().ravel() for i in range(n)]),
               dyn_lines=np.array(span + span2))

    # a14: the NLP the reference's solver worker builds, and its solves (worker loop run for real, warm start carried)
    wk = harness.LmpcSolverWorker()
    tape = wk.solver.tape
    tape.meta.update(N=20, u_lo=-0.4, u_hi=0.4, source="rlmpc2.py:236-491 with run.py:118-126",
                     opts=str(wk.solver.opts))
    tape.save(os.path.join(TAPES, "lmpc_nlp.npz"))
    pr = nlp_probe(tape, rng, 3, 0.2, lambda r: np.concatenate([r.uniform(-.1, .1, 8), r.uniform(-.3, .3, 2),
                                                                   np.clip(1 + .3 * r.standard_normal(34), .01, 1.9), r.uniform(-.1, .1, 8)]))
    out.update({"probe_" + k: v for k, v in pr.items()})

    d = W.lmpc_inputs(B=12, seed=3)
    w_opt, loss = [], []
    for i in range(12):
        w, l = wk.step(d["x0"][i], d["u_prev"][i], d["pvec"][i], d["ref"][i])
        call = wk.solver.calls[-1]
        assert call["status"] == 0, (i, call["kkt"])
        w_opt.append(w); loss.append(l[0])
        print(f"lmpc {i:3d} J={l[0]:.9f} u0={w[168:170]} it={call['iters']} kkt={call['kkt']:.1e}", flush=True)
    rec = solve_record(wk.solver.calls)
    out.update(state=d["x0"][:12], u_prev=d["u_prev"][:12], pvec=d["pvec"][:12], target=d["ref"][:12],
               w=np.array(w_opt), J=np.array(loss), lam_g=rec["lam_g"], iters=rec["iters"], kkt=rec["kkt"], warm=rec["x0"])

    # a15: RLMPC.solve, the reference's own method (rlmpc2.py:986-1021), against the running worker: calls with and
    # without a fresh solution (the "no fresh solution -> shift plan" branch), u_prev fed back through views["control"].
    state = d["x0"][0].copy()
    target = d["ref"][0].copy()
    fac = harness.rlmpc_facade(wk, lambda: state.copy())
    wk.views["model_params"][:] = d["pvec"][0]
    wk.views["control"][:] = 0.0
    wk.events["ctrl_ready"].clear()
    script = [False, True, False, False, True, False, True, True, False, False, False]   # does the worker finish a solve before this call?
    fa = {k: [] for k in ("fresh", "u", "loss", "state", "w_opt", "control_seen")}
    for k, fresh in enumerate(script):
        if fresh:
            wk.run_once()                        # worker solves with what is in shared memory (previous post)
            fa["control_seen"].append(wk.solver.calls[-1]["p"][8:10].copy())
        else:
            fa["control_seen"].append(np.full(2, np.nan))
        u, l = fac.solve(target)
        fa["fresh"].append(fresh); fa["u"].append(u.copy()); fa["loss"].append(np.asarray(l, float).reshape(-1)[0])
        fa["state"].append(state.copy()); fa["w_opt"].append(wk.views["w_opt"].copy())
        # surrogate plant step with the command (run.py applies -u; the surrogate uses the model's own sign)
        state = omodels.lmpc_step(state, u, d["pvec"][0], 0.002)
    out.update({"facade_" + k: np.array(v) for k, v in fa.items()})
    out.update(facade_pvec=d["pvec"][0], facade_target=target)
    wk.close()
    np.savez_compressed(os.path.join(HERE, "ref_lmpc.npz"), **out)


# ----------------------------------------------------------------------------------------------------------------------
def make_policy():
    import glob
    import torch
    rng = np.random.default_rng(14)
    mod = loader.load(loader.LMPC_FILE)
    out = {}
    # a16: Policy as the reference constructs it (rlmpc2.py:33-69), BASELINE config 4's random init
    torch.manual_seed(3)
    pol = mod.Policy(520, 34, {})
    obs = rng.standard_normal((64, 520)).astype(np.float32)
    with torch.no_grad():
        mean, std, value = pol(torch.from_numpy(obs))
    sd = {k: v.detach().numpy().copy() for k, v in pol.state_dict().items()}
    out.update({"init_" + k.replace(".", "__"): v for k, v in sd.items()})
    out.update(init_obs=obs, init_mean=mean.numpy(), init_std=std.numpy(), init_value=value.numpy())

    # the checkpoints the reference ships: mean_net weights + the reference Policy's outputs on seeded observations
    names = []
    for path in sorted(glob.glob(os.path.join(loader.REF_ROOT, "LMPC/src/checkpoints/*/best_agent.pth"))):
        nm = os.path.basename(os.path.dirname(path))
        ck = torch.load(path, map_location="cpu", weights_only=False)
        p2 = mod.Policy(520, 34, {})
        p2.load_state_dict(ck["model"])
        p2.eval()
        o = rng.standard_normal((16, 520)).astype(np.float32)
        with torch.no_grad():
            m, s, v = p2(torch.from_numpy(o))
        names.append(nm)
        for k, t in ck["model"].items():
            if k.startswith("mean_net"):
                out[f"ck_{nm}_{k.replace('.', '__')}"] = t.numpy().copy()
        out[f"ck_{nm}_obs"], out[f"ck_{nm}_mean"], out[f"ck_{nm}_std"], out[f"ck_{nm}_value"] = o, m.numpy(), s.numpy(), v.numpy()
        print("checkpoint", nm, "episode", ck.get("episode"), "return", ck.get("return"))
    out["ck_names"] = np.array(names)

    # compute_gae (rlmpc2.py:589-596) from its own source text
    gae, span = loader.nested_function(loader.LMPC_FILE, ("RLMPC", "_rl_worker"), "compute_gae")
    T = 40
    rew = rng.standard_normal(T).tolist(); val = rng.standard_normal(T).tolist()
    dones = (rng.random(T) < 0.1).astype(float).tolist()
    out.update(gae_rewards=np.array(rew), gae_values=np.array(val), gae_dones=np.array(dones), gae_last=0.37,
               gae_adv=np.array(gae(list(rew), list(val), list(dones), 0.37, 0.99, 0.95)), gae_lines=np.array(span))

    # a17: the RL worker's evaluation loop run for real on the shipped "general" checkpoint: observation build, Welford
    # normaliser, history, policy forward, sampled action, logit-space update every 8th step, smoothed soft-clipped write.
    ckdir = os.path.join(loader.REF_ROOT, "LMPC/src/checkpoints/general")
    rl_packet = {"nx": 8, "nu": 2, "lr": 3e-4, "policy_std_init": 0.1, "clip_eps": 0.2, "epochs": 16, "mini_batch_size": 64,
                 "rollout_len": 2048, "gamma": 0.99, "gae_lambda": 0.95, "obs_dim": 8, "max_param_abs": 2.0,
                 "max_delta_abs": 0.02, "vf_coef": 0.5, "ent_coef": 0.01, "w_pos": 40.0, "w_vel": 0.1, "w_ctrl": 10.0,
                 "max_episode_steps": 20000, "checkpoint_dir": ckdir, "train": False, "seed": 5}       # rlmpc2.py:204-226 from run.py:118-151
    rw = harness.RlWorker(rl_packet, torch_seed=7, numpy_seed=7)
    k0 = rw.views["model_params"].copy()
    T = 34
    st = np.zeros((T, 8)); tg = np.zeros((T, 8)); ct = np.zeros((T, 2)); km = np.zeros((T, 34))
    s = np.zeros(8); s[[0, 2]] = rng.uniform(-0.05, 0.05, 2)
    for t in range(T):
        s = s + 0.002 * rng.standard_normal(8)
        st[t] = s
        tg[t, [0, 2]] = [0.06, -0.04]
        ct[t] = 0.1 * np.sin(0.3 * t + np.array([0.0, 1.0]))
        km[t] = rw.step(st[t], tg[t], ct[t])
    rec = rw.rec
    out.update(rl_k0=k0, rl_state=st, rl_target=tg, rl_control=ct, rl_model_params=km,
               rl_obs=np.concatenate(rec["obs"]), rl_mean=np.concatenate(rec["mean"]), rl_std=rec["std"],
               rl_raw_action=np.concatenate(rec["raw_action"]))
    rw.close()
    print("rl eval trace:", out["rl_obs"].shape, "param drift", np.abs(km[-1] - k0).max())

    # PPO update (rlmpc2.py:775-817) run for real in training mode: tiny rollout so that one update happens.
    import tempfile
    tmp = tempfile.mkdtemp(prefix="dart_ref_ck_")
    tp = dict(rl_packet, train=True, checkpoint_dir=tmp, rollout_len=8, mini_batch_size=4, epochs=2, seed=6)
    rw = harness.RlWorker(tp, torch_seed=9, numpy_seed=9)
    pol = rw.rec["policies"][0]
    before = {k: v.detach().numpy().copy() for k, v in pol.state_dict().items()}
    T = 8 * 8 - 7          # transitions are recorded every 8th step (t = 0, 8, ..., 56): the 8th fills the rollout
    s = np.zeros(8); s[[0, 2]] = rng.uniform(-0.05, 0.05, 2)
    st = np.zeros((T, 8)); tg = np.zeros((T, 8)); ct = np.zeros((T, 2))
    for t in range(T):
        s = s + 0.002 * rng.standard_normal(8)
        st[t] = s; tg[t, [0, 2]] = [0.05, 0.03]; ct[t] = 0.05 * np.cos(0.2 * t + np.array([0.0, 0.5]))
        rw.step(st[t], tg[t], ct[t])
    after = {k: v.detach().numpy().copy() for k, v in pol.state_dict().items()}
    moved = max(np.abs(after[k] - before[k]).max() for k in after)
    print("ppo update: max parameter change", moved)
    assert moved > 0, "the reference's PPO update did not run"
    opt = rw.rec["optimizers"][0].state_dict()
    out.update({"ppo_before_" + k.replace(".", "__"): v for k, v in before.items()})
    out.update({"ppo_after_" + k.replace(".", "__"): v for k, v in after.items()})
    bufrec = rw.rec["buffer"][:8]
    out.update(ppo_buf_obs=np.array([b[1] for b in bufrec]), ppo_buf_act=np.array([b[2] for b in bufrec]),
               ppo_buf_logp=np.array([b[3] for b in bufrec]), ppo_buf_r_slot=np.array([b[4] for b in bufrec]),
               ppo_buf_v_slot=np.array([b[5] for b in bufrec]), ppo_buf_done=np.array([b[6] for b in bufrec]),
               ppo_hparams=np.array([tp["lr"], 1e-5, tp["clip_eps"], tp["vf_coef"], tp["ent_coef"], tp["gamma"], tp["gae_lambda"],
                                     tp["epochs"], tp["mini_batch_size"], 9]))
    out.update(ppo_state=st, ppo_target=tg, ppo_control=ct, ppo_obs=np.concatenate(rw.rec["obs"]),
               ppo_raw_action=np.concatenate(rw.rec["raw_action"]), ppo_value=np.concatenate(rw.rec["value"]),
               ppo_adam_step=np.array([float(opt["state"][i]["step"]) for i in sorted(opt["state"])]),
               ppo_adam_keys=np.array(list(after.keys())))
    rw.close()
    np.savez_compressed(os.path.join(HERE, "ref_policy.npz"), **out)


if __name__ == "__main__":
    if not loader.available():
        sys.exit("needs /root/reference (run in the build container)")
    os.makedirs(TAPES, exist_ok=True)
    which = sys.argv[1:] or ["pmpc", "rmpc", "lmpc", "policy"]
    for w in which:
        {"pmpc": make_pmpc, "rmpc": make_rmpc, "lmpc": make_lmpc, "policy": make_policy}[w]()
    print("done")
