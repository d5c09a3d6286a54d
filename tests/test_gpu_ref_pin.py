"""The CUDA path against what the reference's OWN source computes (fixtures of tests/golden/make_ref_golden.py).

On the GPU box there is no /root/reference: the reference's expression graphs travel as tapes (tests/golden/tapes),
replayed by oracle.tape, so the reference's objective, constraints and their derivatives are evaluated AT THE CUDA
SOLUTIONS.  Bars: north-star |du0| <= 1e-4 rad and |dJ|/J <= 1e-6 against the reference NLP's KKT point; reference
constraint violation <= 1e-8; reference objective at the CUDA point = reported J to 1e-9; KKT stationarity of the
reference NLP (multipliers fitted by least squares) <= 1e-6 relative to the gradient scale.
"""
import ctypes as C
import os

import numpy as np
import pytest

import dart_b200
from oracle import policy
from oracle.tape import Tape
from tests import helpers

pytestmark = pytest.mark.gpu
G = os.path.join(helpers.ROOT, "tests", "golden")
load = lambda n: np.load(os.path.join(G, n), allow_pickle=False)
tape = lambda n: Tape.load(os.path.join(G, "tapes", n))


def reference_kkt_at(tp, w, p, lbx, ubx, lbg=None, ubg=None, act_tol=1e-3):
    """Feasibility and stationarity of the REFERENCE NLP at w.  Multipliers: least squares over [lam_g on equality rows and
    near-active inequality rows; bound multipliers of near-active variable bounds]; returns (violation, f, relative
    stationarity residual, worst of {multiplier sign violation, complementarity |z * slack|} relative to the multiplier
    scale).  ``act_tol`` is wide on purpose: an interior-point solution leaves a weakly active bound at slack ~ mu / z
    (1e-9 / 1e-5 = 1e-4), so its multiplier must be allowed in the fit; a bound that is not really active then gets a
    multiplier ~ 0 and the complementarity term checks exactly that."""
    ev = tp.eval(x=w, p=p)
    g, f = ev["g"], float(ev["f"][0])
    m = len(g)
    lbg = np.zeros(m) if lbg is None else lbg
    ubg = np.zeros(m) if ubg is None else ubg
    viol = max(np.maximum(lbg - g, 0).max(), np.maximum(g - ubg, 0).max(), np.maximum(lbx - w, 0).max(), np.maximum(w - ubx, 0).max())
    gf = tp.grad("f", "x", x=w, p=p)
    Jg = tp.jac("g", "x", x=w, p=p)
    eq = lbg == ubg
    act_hi = ~eq & (ubg - g <= act_tol)
    act_lo = ~eq & (g - lbg <= act_tol)
    rows = eq | act_hi | act_lo
    at_lo, at_hi = w - lbx <= act_tol, ubx - w <= act_tol
    bnd = np.nonzero(at_lo | at_hi)[0]
    A = np.concatenate([Jg[rows].T, np.eye(len(w))[:, bnd]], axis=1)
    sol, *_ = np.linalg.lstsq(A, -gf, rcond=None)
    res = np.abs(A @ sol + gf).max() / max(1.0, np.abs(gf).max())
    lam = np.zeros(m); lam[rows] = sol[:rows.sum()]
    z = sol[rows.sum():]
    sign = 0.0
    scale = max(1.0, np.abs(sol).max())
    if act_hi.any():
        sign = max(sign, (-lam[act_hi]).max() / scale)
    if act_lo.any():
        sign = max(sign, (lam[act_lo]).max() / scale)
    for j, b in enumerate(bnd):                      # z = z_U - z_L: >= 0 at an upper bound, <= 0 at a lower bound
        sign = max(sign, (-z[j] if at_hi[b] else z[j]) / scale)
        sign = max(sign, abs(z[j]) * (ubx[b] - w[b] if at_hi[b] else w[b] - lbx[b]) / scale)
    if act_hi.any():
        sign = max(sign, (np.abs(lam[act_hi]) * (ubg - g)[act_hi]).max() / scale)
    if act_lo.any():
        sign = max(sign, (np.abs(lam[act_lo]) * (g - lbg)[act_lo]).max() / scale)
    return viol, f, res, sign


def check_against_reference(out, d, u0_ref, tapes_for, p_for, bounds_for, what):
    assert (out["status"] == 0).all(), (what, out["status"])
    du0 = np.abs(out["u0"] - u0_ref).max()
    dJ = (np.abs(out["J"] - d["J"]) / np.abs(d["J"])).max()
    dw = np.abs(out["w"] - d["w"]).max()
    assert du0 <= helpers.TOL_U0 and dJ <= helpers.TOL_J, (what, du0, dJ)
    worst = [0.0, 0.0, 0.0, 0.0]
    for i in range(len(d["J"])):
        tp = tapes_for(i)
        viol, f, res, sign = reference_kkt_at(tp, out["w"][i], p_for(i), *bounds_for(tp))
        worst = [max(worst[0], viol), max(worst[1], abs(f - out["J"][i]) / abs(f)), max(worst[2], res), max(worst[3], sign)]
    print(f"{what}: vs reference solve |du0|={du0:.2e} rel dJ={dJ:.2e} |dw|={dw:.2e}; reference NLP at the CUDA point: "
          f"violation {worst[0]:.2e}, |f-J|/f {worst[1]:.2e}, stationarity {worst[2]:.2e}, sign {worst[3]:.2e}")
    assert worst[0] <= 1e-8 and worst[1] <= 1e-9 and worst[2] <= 1e-6 and worst[3] <= 1e-6, (what, worst)


def test_pmpc_cuda_solutions_on_the_reference_nlp(built):
    d = load("ref_pmpc.npz")
    eng = dart_b200.NMPCEngine(dart_b200.pmpc_cfg(), device=0)
    aux = np.stack([d["Qp"], d["Qv"], d["R"], d["mu"]], axis=1)
    out = eng.solve(d["state"], d["target"], aux=aux)
    tapes = {c: tape(f"pmpc_nlp_{c}.npz") for c in np.unique(d["combo"])}
    check_against_reference(out, d, d["u0"], lambda i: tapes[int(d["combo"][i])],
                            lambda i: np.concatenate([d["state"][i], d["target"][i]]),
                            lambda tp: (tp.meta["lbx"], tp.meta["ubx"]), "pmpc config 2 sample")
    # config 1 through the drop-in class (PMPC/main.py:59-69 parameters), against the reference class' own return values
    model = dart_b200.GravityModel(-9.81, 0.002); data = dart_b200.StateHolder()
    mpc = dart_b200.PMPC(model, data, 0.002, nx=6, nu=2, N=15, Qp=400, Qv=2, R=0.2, mu=0.10, u_bounds=(-0.6, 0.6))
    b = data.body(mpc.target_body); b.xpos[:] = [0, 0, 0.43]; b.cvel[3:6] = 0
    u, loss = mpc.solve(np.array([0.1, 0, 0.05, 0, 0.4, 0]))
    assert np.abs(u - d["c1_u0"]).max() <= helpers.TOL_U0 and abs(loss[0] - d["c1_J"][0]) <= helpers.TOL_J * d["c1_J"][0]
    tp = tape("pmpc_nlp_config1.npz")
    viol, f, res, sign = reference_kkt_at(tp, mpc.w0, np.array([0, 0, 0, 0, .43, 0, .1, 0, .05, 0, .4, 0]), tp.meta["lbx"], tp.meta["ubx"])
    assert viol <= 1e-8 and abs(f - loss[0]) <= 1e-9 * f and res <= 1e-6 and sign <= 1e-6


def test_rmpc_cuda_solutions_on_the_reference_nlp(built):
    d = load("ref_rmpc.npz")
    tp = tape("rmpc_nlp.npz")
    eng = dart_b200.NMPCEngine(dart_b200.rmpc_cfg(), device=0)
    out = eng.solve(d["x0"], d["ref"], aux=np.concatenate([d["u_prev"], d["theta"]], axis=1))
    check_against_reference(out, d, d["u0"], lambda i: tp,
                            lambda i: np.concatenate([d["x0"][i], d["u_prev"][i], d["theta"][i], d["ref"][i]]),
                            lambda t: (t.meta["lbx"], t.meta["ubx"], t.meta["lbg"], t.meta["ubg"]), "rmpc")


def test_lmpc_cuda_solutions_on_the_reference_nlp(built):
    d = load("ref_lmpc.npz")
    tp = tape("lmpc_nlp.npz")
    eng = dart_b200.NMPCEngine(dart_b200.lmpc_cfg(), device=0)
    out = eng.solve(d["state"], d["target"], aux=np.concatenate([d["u_prev"], d["pvec"]], axis=1))
    lbx = np.concatenate([np.full(168, -np.inf), np.full(40, -0.4)])
    check_against_reference(out, d, d["w"][:, 168:170], lambda i: tp,
                            lambda i: np.concatenate([d["state"][i], d["u_prev"][i], d["pvec"][i], d["target"][i]]),
                            lambda t: (lbx, -lbx), "lmpc")


def test_rls_kernel_matches_reference_class_trajectory(built):
    import torch
    d = load("ref_rmpc.npz")
    dev = torch.device("cuda", 0)
    E, T = d["rls_y"].shape
    th = torch.zeros((E, 1, 7), dtype=torch.float64, device=dev)
    P = (torch.eye(7, dtype=torch.float64, device=dev) * 1e3).repeat(E, 1, 1, 1).contiguous()
    worst = 0.0
    for t in range(T):
        dart_b200.rls_update_device(th, P, torch.from_numpy(d["rls_phi"][:, t].copy()).to(dev), torch.from_numpy(d["rls_y"][:, t, None].copy()).to(dev), 0.995)
        ref = d["rls_theta"][:, t]
        worst = max(worst, np.abs(th.cpu().numpy()[:, 0] - ref).max() / max(1e-12, np.abs(ref).max()))
    assert worst <= helpers.TOL_RLS, worst
    assert np.abs(P.cpu().numpy()[:, 0] - d["rls_P"][:, -1]).max() <= 1e-6 * np.abs(d["rls_P"][:, -1]).max()
    # the host drop-in class
    r = dart_b200.RLS(p=7, theta0=np.zeros(7), P0=1e3, lam=0.995)
    for t in range(T):
        r.update(d["rls_phi"][0, t], d["rls_y"][0, t])
    assert np.abs(r.get() - d["rls_theta"][0, -1]).max() <= 1e-9 * np.abs(d["rls_theta"][0, -1]).max()


def test_rmpc_closed_loop_matches_the_reference_main_loop(built):
    """RMPCBatch (RLS + governor + staged reference + warm-started solve, all on the device) against the fixture produced
    by the reference's classes inside the reference's loop body (rob_ctrl.py:330-352)."""
    import torch
    d = load("ref_rmpc.npz")
    B, T = d["loop_u0"].shape[:2]
    x = d["loop_x0"].copy(); x[:, [1, 3]] *= 0.5
    dev = torch.device("cuda", 0)
    ctl = dart_b200.RMPCBatch(B, d["loop_target"], x, device=0, rls_P0=1.0)
    rv0 = np.zeros((B, 4)); rv0[:, [0, 2]] = x[:, [0, 2]]
    ctl.set_virtual_reference(rv0)
    assert np.array_equal(dart_b200.AdaptiveNPMPCSmooth.build_ref_traj(None, d["ref_rv"][0], d["ref_target"][0], 20, 4, 0.2), d["ref_out"][0])
    for t in range(T):
        assert np.abs(x - d["loop_x"][:, t]).max() <= 1e-7
        u = ctl.step(torch.from_numpy(x).to(dev)).cpu().numpy()
        assert np.abs(u - d["loop_u0"][:, t]).max() <= helpers.TOL_U0, (t, np.abs(u - d["loop_u0"][:, t]).max())
        th = ctl.theta.cpu().numpy().reshape(B, 14)
        assert np.abs(th - d["loop_theta"][:, t]).max() <= helpers.TOL_RLS * max(1e-9, np.abs(d["loop_theta"][:, t]).max())
        assert (np.abs(ctl.J.cpu().numpy() - d["loop_J"][:, t]) / d["loop_J"][:, t]).max() <= helpers.TOL_J
        x = dart_b200.workloads.rmpc_plant_step(x, d["loop_u0"][:, t], d["loop_mu_plant"], d["loop_c_plant"])
    assert (ctl.status.cpu().numpy() == 0).all()


def _mean_net(d, prefix):
    return [(d[f"{prefix}mean_net__{i}__weight"], d[f"{prefix}mean_net__{i}__bias"]) for i in (0, 2, 4)]


TOL_MLP_REF = 2e-5          # FP32-fidelity forward (see test_gpu_lmpc_policy.TOL_MLP)


def test_policy_kernel_matches_reference_policy_class(built):
    """``Policy.forward`` of the reference on its random init (BASELINE config 4) and on the nine checkpoints it ships."""
    import torch
    d = load("ref_policy.npz")
    tol = getattr(helpers, "TOL_MLP", TOL_MLP_REF)
    mine = dart_b200.init_policy_weights(seed=3)
    for (W, b), (Wr, br) in zip(mine, _mean_net(d, "init_")):
        assert np.array_equal(W, Wr) and np.array_equal(b, br)          # same random init as Policy(520, 34, {}) under seed 3
    pol = dart_b200.PolicyMLP(mine, device=0)
    out = pol.forward(torch.from_numpy(d["init_obs"]).cuda()).cpu().numpy()
    assert np.abs(out - d["init_mean"]).max() <= tol
    pol.close()
    for nm in d["ck_names"]:
        pol = dart_b200.PolicyMLP(_mean_net(d, f"ck_{nm}_"), device=0)
        out = pol.forward(torch.from_numpy(d[f"ck_{nm}_obs"]).cuda()).cpu().numpy()
        err = np.abs(out - d[f"ck_{nm}_mean"]).max()
        assert err <= tol * max(1.0, np.abs(d[f"ck_{nm}_mean"]).max()), (nm, err)
        pol.close()


def test_lmpc_batch_replays_the_reference_rl_worker(built):
    """LMPCBatch (obs push -> policy -> parameter update) fed the inputs and the sampled actions of the reference's RL worker
    run (evaluation mode, shipped checkpoint): observations, policy means and the shared-memory parameters step by step."""
    import torch
    from tests.test_ref_pin import reference_initial_current_k
    d = load("ref_policy.npz")
    tol = getattr(helpers, "TOL_MLP", TOL_MLP_REF)
    dev = torch.device("cuda", 0)
    cur_k = reference_initial_current_k(5)
    ctl = dart_b200.LMPCBatch(1, d["rl_k0"][None], obs_k0=cur_k[None], weights=_mean_net(d, "ck_general_"), device=0)
    seen = {}

    def source(obs, action_out):
        t = seen["t"]
        o = obs.cpu().numpy()
        assert np.abs(o - d["rl_obs"][t]).max() <= 2e-6 * max(1.0, np.abs(d["rl_obs"][t]).max()), t
        ctl.policy.forward(obs, action_out)
        assert np.abs(action_out.cpu().numpy() - d["rl_mean"][t]).max() <= tol * max(1.0, np.abs(d["rl_mean"][t]).max())
        action_out.copy_(torch.from_numpy(d["rl_raw_action"][t:t + 1]).to(dev))       # the reference applied its SAMPLE (rlmpc2.py:678)

    ctl.action_source = source
    t64 = lambda a: torch.from_numpy(np.ascontiguousarray(a[None])).to(dev)
    for t in range(len(d["rl_state"])):
        seen["t"] = t
        ctl.u_prev.copy_(t64(d["rl_control"][t]))              # views["control"] as the worker read it
        ctl.step(t64(d["rl_state"][t]), t64(d["rl_target"][t]))
        k = ctl.pvec.cpu().numpy()[0]
        assert np.abs(k - d["rl_model_params"][t]).max() <= 1e-6, (t, np.abs(k - d["rl_model_params"][t]).max())


def test_rlmpc_facade_plan_shift_matches_reference_solve(built):
    """``RLMPC.solve``'s "no fresh solution" branch (rlmpc2.py:1013-1018): the command is the next entry of the last plan.
    The reference fixture interleaves fresh and stale calls; LMPCBatch(plan_fallback) is driven with the same pattern by
    marking the stale solves unusable, and must publish the same commands as the reference's own ``solve`` did."""
    import torch
    d = load("ref_lmpc.npz")
    fresh, u_ref = d["facade_fresh"], d["facade_u"]
    dev = torch.device("cuda", 0)
    ctl = dart_b200.LMPCBatch(1, d["facade_pvec"][None], device=0, plan_fallback=True, update_every=0, warm_mu=None)
    t64 = lambda a: torch.from_numpy(np.ascontiguousarray(a[None])).to(dev)
    n_stale = 0
    for i in range(len(fresh)):
        # a fresh plan was computed from the state posted by the PREVIOUS call (the worker is one mailbox exchange behind)
        state = d["facade_state"][i - 1] if fresh[i] else d["facade_state"][i]
        u = ctl.step(t64(state), t64(d["facade_target"]), fresh=torch.tensor([bool(fresh[i])], device=dev)).cpu().numpy()[0]
        assert np.abs(u - u_ref[i]).max() <= helpers.TOL_U0, (i, u, u_ref[i])
        if fresh[i]:
            assert np.abs(ctl.w.cpu().numpy()[0] - d["facade_w_opt"][i]).max() <= 1e-5
            assert np.abs(ctl.aux.cpu().numpy()[0, :2] - u_ref[i]).max() <= helpers.TOL_U0      # next u_prev = published command
        else:
            n_stale += 1
    assert int(ctl.n_fallback) == n_stale and n_stale >= 5


def test_ppo_update_kernel_matches_reference_worker_update(built):
    """The reference's RL worker ran one real PPO update (fixture); the CUDA learner from the same parameters, buffer and
    minibatch order must land on the same parameters (Adam steps of lr = 3e-4: bar 2 % of one step per element for 99 %
    of the elements, see DESIGN 8d for why not all)."""
    import torch
    from oracle import ppo as oppo
    d = load("ref_policy.npz")
    lr, wd, clip_eps, vf, ent, gamma, lam, epochs, mb, npseed = d["ppo_hparams"]
    keys = [k[len("ppo_before_"):] for k in d.files if k.startswith("ppo_before_")]
    sd0 = {k.replace("__", "."): d["ppo_before_" + k] for k in keys}
    sd1 = {k.replace("__", "."): d["ppo_after_" + k] for k in keys}
    tr = dart_b200.PPOTrainer(capacity=64, state_dict=sd0, lr=float(lr), weight_decay=float(wd), clip_eps=float(clip_eps),
                              vf_coef=float(vf), ent_coef=float(ent), epochs=int(epochs), mini_batch_size=int(mb),
                              gamma=float(gamma), gae_lambda=float(lam))
    dev = tr.dev
    f32 = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(dev)
    obs, act, logp = f32(d["ppo_buf_obs"]), f32(d["ppo_buf_act"]), f32(d["ppo_buf_logp"])
    # rollout-time quantities first: the learner's own forward reproduces the stored log-probabilities and values
    _, logp_dev, val_dev, _ = tr.act(obs, None)
    rewards, values, dones = d["ppo_buf_r_slot"], d["ppo_buf_v_slot"], d["ppo_buf_done"]
    last_val = tr.act(obs[-1:].contiguous(), None)[2]
    T = len(rewards)
    adv, ret = tr.gae(f32(rewards).view(T, 1), f32(values).view(T, 1), f32(dones).view(T, 1), last_val.view(1))
    adv_ref = np.array(oppo.compute_gae(rewards.tolist(), values.tolist(), dones.tolist(), float(last_val[0]), float(gamma), float(lam)))
    assert np.abs(adv.cpu().numpy()[:, 0] - adv_ref).max() <= 1e-5 * max(1.0, np.abs(adv_ref).max())
    tr.normalize_(ret, 0); tr.normalize_(adv, 1)
    flat = (obs, act, logp, adv.reshape(T).contiguous(), ret.reshape(T).contiguous())
    np.random.seed(int(npseed))
    for _ in range(int(epochs)):
        idxs = np.random.permutation(T)
        for s in range(0, T, int(mb)):
            tr.update_minibatch(*flat, idx=torch.from_numpy(idxs[s:s + int(mb)]).to(dev))
    got = tr.state_dict()
    frac_ok, worst, n = 0, 0.0, 0
    for k in got:
        e = np.abs(got[k] - sd1[k])
        moved = np.abs(sd1[k] - sd0[k])
        worst = max(worst, e.max())
        frac_ok += (e <= 0.02 * float(lr) * int(epochs) * (T // int(mb)) + 1e-9).sum()
        n += e.size
        assert e.max() <= 1.01 * moved.max() + 1e-7, k
    print(f"ppo update vs reference run: worst |dparam| = {worst:.2e} (lr {lr:g}), within 2% of the steps: {frac_ok / n:.4f}")
    assert frac_ok / n >= 0.99 and worst <= float(lr)
    tr.close()
