"""The oracle against what the reference's OWN source computes (CPU; no GPU, no /root/reference needed).

``tests/golden/ref_*.npz`` and ``tests/golden/tapes/*.npz`` were written by ``tests/golden/make_ref_golden.py`` by
executing the reference's unmodified files through ``oracle/refshim`` (see there).  Bars: 1e-12 relative for float64
function values (the same arithmetic restated), 1e-9 for derivatives (complex step on both sides), 1e-6 for float32.
Solutions: first move within 1e-6 rad and objective within 1e-9 relative of the reference NLP's KKT point (two different
solvers, both at 1e-10), far inside the product bars of BASELINE.json (1e-4 rad / 1e-6).
"""
import os

import numpy as np
import pytest

from oracle import ipm, models, nlp_layout, policy, problems, rls
from oracle.tape import Tape

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return np.load(os.path.join(G, name), allow_pickle=False)


def tape(name):
    return Tape.load(os.path.join(G, "tapes", name))


def rel(a, b):
    return np.abs(a - b).max() / max(1e-300, np.abs(b).max())


def kkt_residual(tp, w, p, lam_g, lbx, ubx, lbg=None, ubg=None, act_tol=1e-7):
    """Stationarity / feasibility / sign conditions of a stored solution on the reference's own graph."""
    ev = tp.eval(x=w, p=p)
    g = ev["g"]
    lbg = np.zeros_like(g) if lbg is None else lbg
    ubg = np.zeros_like(g) if ubg is None else ubg
    feas = max(np.maximum(lbg - g, 0).max(), np.maximum(g - ubg, 0).max(), np.maximum(lbx - w, 0).max(), np.maximum(w - ubx, 0).max())
    r = tp.grad("f", "x", x=w, p=p) + tp.jac("g", "x", x=w, p=p).T @ lam_g
    at_lo, at_hi = w - lbx <= act_tol, ubx - w <= act_tol
    free = ~(at_lo | at_hi)
    stat = np.abs(r[free]).max()
    sign = max(0.0, (-r[at_lo]).max() if at_lo.any() else 0.0, (r[at_hi]).max() if at_hi.any() else 0.0)   # r = z_L - z_U
    # inequality rows: multiplier sign and complementarity
    ineq = lbg < ubg
    comp = 0.0
    if ineq.any():
        lam = lam_g[ineq]
        slack_hi, slack_lo = (ubg - g)[ineq], (g - lbg)[ineq]
        comp = max(np.abs(np.where(lam > 0, lam * np.where(np.isfinite(slack_hi), slack_hi, 0.0), 0.0)).max(),
                   np.abs(np.where(lam < 0, lam * np.where(np.isfinite(slack_lo), slack_lo, 0.0), 0.0)).max())
    return feas, stat, sign, comp


# ------------------------------------------------------------------------------------------------------------------ PMPC
def test_pmpc_dynamics_match_reference_function():
    d = load("ref_pmpc.npz")
    out = models.pmpc_step(d["dyn_x"], d["dyn_u"], -9.81, d["dyn_mu"], 0.002)
    assert rel(out, d["dyn_out"]) <= 1e-12


def test_pmpc_nlp_functions_match_reference_graph():
    d = load("ref_pmpc.npz")
    for k in range(len(d["probe_f"])):
        Qp, Qv, R, mu = d["combo_params"][int(d["probe_combo"][k])]
        w, p = d["probe_w"][k], d["probe_p"][k]
        f, g = nlp_layout.pmpc_nlp(w, p, Qp, Qv, R, mu)
        assert abs(f - d["probe_f"][k]) <= 1e-12 * abs(d["probe_f"][k])
        assert np.abs(g - d["probe_g"][k]).max() <= 1e-12 * np.abs(d["probe_g"][k]).max()
        gf, jg = nlp_layout.complex_step(lambda wc: nlp_layout.pmpc_nlp(wc, p, Qp, Qv, R, mu), w)
        assert rel(gf, d["probe_gradf"][k]) <= 1e-9
        assert np.abs(jg - d["probe_jg"][k]).max() <= 1e-9 * np.abs(d["probe_jg"][k]).max()
        # and the stored numbers are what the travelling tape replays
        tp = tape(f"pmpc_nlp_{int(d['probe_combo'][k])}.npz")
        ev = tp.eval(x=w, p=p)
        assert ev["f"][0] == d["probe_f"][k] and np.array_equal(ev["g"], d["probe_g"][k])


def test_pmpc_reference_solutions_are_kkt_points_of_the_reference_graph():
    d = load("ref_pmpc.npz")
    for i in range(len(d["J"])):
        tp = tape(f"pmpc_nlp_{int(d['combo'][i])}.npz")
        p = np.concatenate([d["state"][i], d["target"][i]])
        feas, stat, sign, _ = kkt_residual(tp, d["w"][i], p, d["lam_g"][i], tp.meta["lbx"], tp.meta["ubx"])
        assert feas <= 1e-9 and stat <= 1e-7 and sign <= 1e-7, (i, feas, stat, sign)
        assert abs(tp.eval(x=d["w"][i], p=p)["f"][0] - d["J"][i]) <= 1e-12 * abs(d["J"][i])


def test_pmpc_oracle_solutions_match_reference_solve():
    d = load("ref_pmpc.npz")
    prob = problems.pmpc_problem(d["state"], d["target"], Qp=d["Qp"], Qv=d["Qv"], R=d["R"], mu=d["mu"])
    sol = ipm.solve(prob, opts=ipm.Options(tol=1e-10))
    assert (sol["status"] == 0).all()
    assert np.abs(sol["U"][:, 0] - d["u0"]).max() <= 1e-6
    assert (np.abs(sol["J"] - d["J"]) / np.abs(d["J"])).max() <= 1e-9
    w = problems.pack_w(prob, problems.pmpc_full_states(prob, sol["U"]), sol["U"])        # reference layout incl. z rows
    Xz = problems.pmpc_full_states(prob, sol["U"])
    Xz[:, :, :4] = sol["X"]
    w = np.concatenate([Xz.reshape(len(w), -1), sol["U"].reshape(len(w), -1)], axis=1)
    assert np.abs(w - d["w"]).max() <= 1e-6
    # config 1 (PMPC/main.py weights, README example)
    c1 = problems.pmpc_problem([[0, 0, 0, 0, .43, 0]], [[.1, 0, .05, 0, .4, 0]], Qp=400.0, Qv=2.0, R=0.2, mu=0.1)
    s1 = ipm.solve(c1, opts=ipm.Options(tol=1e-10))
    assert np.abs(s1["U"][0, 0] - d["c1_u0"]).max() <= 1e-6 and abs(s1["J"][0] - d["c1_J"][0]) <= 1e-9 * d["c1_J"][0]


# ------------------------------------------------------------------------------------------------------------------ RMPC
def test_rls_matches_reference_class_trajectory():
    d = load("ref_rmpc.npz")
    E, T = d["rls_y"].shape
    for e in range(E):
        r = rls.RLS(7, np.zeros(7), 1e3, 0.995)
        for t in range(T):
            r.update(d["rls_phi"][e, t], d["rls_y"][e, t])
            assert rel(r.get(), d["rls_theta"][e, t]) <= 1e-12, (e, t)
        assert rel(r.P, d["rls_P"][e, -1]) <= 1e-12
    # batched form used by the closed-loop oracle
    th = np.zeros((E, 1, 7)); P = np.tile(np.eye(7) * 1e3, (E, 1, 1, 1))
    for t in range(T):
        th, P = rls.rls_update_batch(th, P, d["rls_phi"][:, t], d["rls_y"][:, t, None], 0.995)
    assert rel(th[:, 0], d["rls_theta"][:, -1]) <= 1e-12


def test_build_ref_traj_matches_reference_staticmethod():
    d = load("ref_rmpc.npz")
    out = problems.build_ref_traj(None, d["ref_rv"], d["ref_target"], 20, 4, 0.2)
    assert np.array_equal(out, d["ref_out"])


def test_rmpc_dynamics_match_reference_function():
    d = load("ref_rmpc.npz")
    out = models.rmpc_step(d["dyn_x"], d["dyn_u"], d["dyn_th"], -9.81, 0.1, 0.002)
    assert rel(out, d["dyn_out"]) <= 1e-12


def test_rmpc_nlp_functions_match_reference_graph():
    d = load("ref_rmpc.npz")
    tp = tape("rmpc_nlp.npz")
    lo, hi = nlp_layout.rmpc_g_bounds()
    assert np.array_equal(lo, tp.meta["lbg"]) and np.array_equal(hi, tp.meta["ubg"])
    for k in range(len(d["probe_f"])):
        w, p = d["probe_w"][k], d["probe_p"][k]
        f, g = nlp_layout.rmpc_nlp(w, p)
        assert abs(f - d["probe_f"][k]) <= 1e-12 * abs(d["probe_f"][k])
        assert np.abs(g - d["probe_g"][k]).max() <= 1e-12 * np.abs(d["probe_g"][k]).max()
        gf, jg = nlp_layout.complex_step(lambda wc: nlp_layout.rmpc_nlp(wc, p), w)
        assert rel(gf, d["probe_gradf"][k]) <= 1e-9
        assert np.abs(jg - d["probe_jg"][k]).max() <= 1e-9 * np.abs(d["probe_jg"][k]).max()
        assert np.array_equal(tp.eval(x=w, p=p)["g"], d["probe_g"][k])


def test_rmpc_reference_solutions_are_kkt_points_and_oracle_matches():
    d = load("ref_rmpc.npz")
    tp = tape("rmpc_nlp.npz")
    for i in range(len(d["J"])):
        p = np.concatenate([d["x0"][i], d["u_prev"][i], d["theta"][i], d["ref"][i]])
        feas, stat, sign, comp = kkt_residual(tp, d["w"][i], p, d["lam_g"][i], tp.meta["lbx"], tp.meta["ubx"], tp.meta["lbg"], tp.meta["ubg"])
        assert feas <= 1e-9 and stat <= 1e-7 and sign <= 1e-7 and comp <= 1e-7, (i, feas, stat, sign, comp)
    sol = ipm.solve(problems.rmpc_problem(d["x0"], d["u_prev"], d["theta"], d["ref"]), opts=ipm.Options(tol=1e-10))
    assert (sol["status"] == 0).all()
    assert np.abs(sol["U"][:, 0] - d["u0"]).max() <= 1e-6
    assert (np.abs(sol["J"] - d["J"]) / np.abs(d["J"])).max() <= 1e-9


def test_rmpc_closed_loop_matches_reference_classes_in_the_reference_main_loop():
    """rob_ctrl.py:330-352 around the reference's RLS / AdaptiveNPMPCSmooth (fixture) vs the oracle's batched loop."""
    import dart_b200
    d = load("ref_rmpc.npz")
    B, T = d["loop_u0"].shape[:2]
    x = d["loop_x0"].copy(); x[:, [1, 3]] *= 0.5
    th = np.zeros((B, 2, 7)); P = np.tile(np.eye(7), (B, 2, 1, 1))
    r_v = np.zeros((B, 4)); r_v[:, [0, 2]] = x[:, [0, 2]]
    prev = x.copy(); u_prev = np.zeros((B, 2)); Xw = np.zeros((B, 21, 6)); Uw = np.zeros((B, 20, 2))
    for t in range(T):
        assert np.abs(x - d["loop_x"][:, t]).max() <= 1e-8
        th, P = rls.rls_update_batch(th, P, rls.regressor(prev, 0.1), rls.accel_measurement(x, prev, 0.002), 0.995)
        assert rel(th.reshape(B, 14), d["loop_theta"][:, t]) <= 1e-6
        r_v = problems.reference_governor(r_v, d["loop_target"])
        assert np.abs(r_v - d["loop_r_v"][:, t]).max() <= 1e-15
        ref = problems.build_ref_traj(x, r_v, d["loop_target"], 20, 4, 0.2)
        sol = ipm.solve(problems.rmpc_problem(x, u_prev, th.reshape(B, 14), ref), X0=Xw, U0=Uw, opts=ipm.Options(tol=1e-10))
        assert (sol["status"] == 0).all()
        Xw, Uw = sol["X"], sol["U"]
        u = sol["U"][:, 0]
        assert np.abs(u - d["loop_u0"][:, t]).max() <= 1e-6, (t, np.abs(u - d["loop_u0"][:, t]).max())
        assert (np.abs(sol["J"] - d["loop_J"][:, t]) / np.abs(d["loop_J"][:, t])).max() <= 1e-8
        prev, u_prev = x.copy(), u.copy()
        x = dart_b200.workloads.rmpc_plant_step(x, u, d["loop_mu_plant"], d["loop_c_plant"])


# ------------------------------------------------------------------------------------------------------------------ LMPC
def test_lmpc_dynamics_match_reference_safe_dynamics():
    d = load("ref_lmpc.npz")
    assert rel(models.lmpc_dynamics(d["dyn_x"], d["dyn_u"], d["dyn_p"]), d["dyn_cont"]) <= 1e-12
    assert rel(models.lmpc_step(d["dyn_x"], d["dyn_u"], d["dyn_p"], 0.002), d["dyn_out"]) <= 1e-12
    assert (d["dyn_p"] < 0).any()          # the |p| squash branch is exercised


def test_lmpc_nlp_functions_match_reference_graph():
    d = load("ref_lmpc.npz")
    tp = tape("lmpc_nlp.npz")
    for k in range(len(d["probe_f"])):
        w, p = d["probe_w"][k], d["probe_p"][k]
        f, g = nlp_layout.lmpc_nlp(w, p)
        assert abs(f - d["probe_f"][k]) <= 1e-12 * abs(d["probe_f"][k])
        assert np.abs(g - d["probe_g"][k]).max() <= 1e-12 * np.abs(d["probe_g"][k]).max()
        gf, jg = nlp_layout.complex_step(lambda wc: nlp_layout.lmpc_nlp(wc, p), w)
        assert rel(gf, d["probe_gradf"][k]) <= 1e-9
        assert np.abs(jg - d["probe_jg"][k]).max() <= 1e-9 * np.abs(d["probe_jg"][k]).max()
        assert np.array_equal(tp.eval(x=w, p=p)["g"], d["probe_g"][k])


def test_lmpc_reference_worker_solutions_are_kkt_points_and_oracle_matches():
    d = load("ref_lmpc.npz")
    tp = tape("lmpc_nlp.npz")
    lbx = np.concatenate([np.full(168, -np.inf), np.full(40, -0.4)]); ubx = -lbx
    for i in range(len(d["J"])):
        p = np.concatenate([d["state"][i], d["u_prev"][i], d["pvec"][i], d["target"][i]])
        feas, stat, sign, _ = kkt_residual(tp, d["w"][i], p, d["lam_g"][i], lbx, ubx)
        assert feas <= 1e-9 and stat <= 1e-7 and sign <= 1e-7, (i, feas, stat, sign)
        if i:
            assert np.array_equal(d["warm"][i], d["w"][i - 1])          # the worker's warm start is the previous w_opt (:519)
    sol = ipm.solve(problems.lmpc_problem(d["state"], d["u_prev"], d["pvec"], d["target"]), opts=ipm.Options(tol=1e-10))
    assert (sol["status"] == 0).all()
    assert np.abs(sol["U"][:, 0] - d["w"][:, 168:170]).max() <= 1e-6
    assert (np.abs(sol["J"] - d["J"]) / np.abs(d["J"])).max() <= 1e-9


def test_rlmpc_solve_mailbox_semantics_of_the_reference():
    """What ``RLMPC.solve`` (rlmpc2.py:986-1021) returned call by call: fresh plan -> U_opt[0]; none -> next entry of the
    previous plan (the list shrinks); the very first call (no plan yet) holds last_control = 0."""
    d = load("ref_lmpc.npz")
    fresh, u, wopt = d["facade_fresh"], d["facade_u"], d["facade_w_opt"]
    assert not fresh[0] and np.array_equal(u[0], [0.0, 0.0])
    plan, k = None, 0
    for i in range(len(fresh)):
        if fresh[i]:
            plan, k = wopt[i][168:].reshape(20, 2), 0
            assert np.array_equal(u[i], plan[0])
        elif plan is not None:
            k += 1
            assert np.array_equal(u[i], plan[k])
    # the worker saw the previous call's command as u_prev (views["control"], :505,:1019)
    seen = d["facade_control_seen"]
    for i in range(1, len(fresh)):
        if fresh[i]:
            assert np.array_equal(seen[i], u[i - 1])


# ------------------------------------------------------------------------------------------------------------------ policy
def _mean_net(d, prefix):
    return [(d[f"{prefix}mean_net__{i}__weight"], d[f"{prefix}mean_net__{i}__bias"]) for i in (0, 2, 4)]


def test_policy_forward_matches_reference_class():
    d = load("ref_policy.npz")
    out = policy.mlp_forward(d["init_obs"], _mean_net(d, "init_"))
    assert np.abs(out - d["init_mean"]).max() <= 1e-6
    assert np.allclose(d["init_std"], 0.1, atol=1e-7)
    for nm in d["ck_names"]:
        out = policy.mlp_forward(d[f"ck_{nm}_obs"], _mean_net(d, f"ck_{nm}_"))
        assert np.abs(out - d[f"ck_{nm}_mean"]).max() <= 2e-6 * max(1.0, np.abs(d[f"ck_{nm}_mean"]).max()), nm
    assert len(d["ck_names"]) == 9


def test_policy_random_init_is_the_reference_construction():
    """BASELINE config 4: ``Policy(520, 34, {})`` under ``torch.manual_seed(3)``."""
    from oracle import ppo
    d = load("ref_policy.npz")
    pol = ppo.make_policy(seed=3)
    for k, v in pol.state_dict().items():
        assert np.array_equal(v.numpy(), d["init_" + k.replace(".", "__")]), k


def test_compute_gae_matches_reference_function():
    from oracle import ppo
    d = load("ref_policy.npz")
    adv = ppo.compute_gae(d["gae_rewards"].tolist(), d["gae_values"].tolist(), d["gae_dones"].tolist(), float(d["gae_last"]), 0.99, 0.95)
    assert np.array_equal(np.array(adv), d["gae_adv"])


def reference_initial_current_k(seed, k_max=2.0, min_k=1e-2, act_dim=34):
    """rlmpc2.py:618-623."""
    rng = np.random.default_rng(seed)
    jitter = rng.uniform(-0.05, 0.05, size=act_dim) * k_max
    return np.clip(np.full(act_dim, 0.5 * k_max) + jitter, min_k, k_max - max(1e-3, 0.05 * k_max))


def test_observation_and_parameter_update_match_reference_rl_worker_trace():
    """The reference's RL worker ran 34 steps on its shipped checkpoint (fixture); replay its inputs through the oracle."""
    d = load("ref_policy.npz")
    cur_k = reference_initial_current_k(5)
    shm = policy.write_params(cur_k, np.zeros(34))                     # write_params_to_shm(current_k) at start-up (:623), shm starts at 0
    assert rel(shm, d["rl_k0"]) <= 1e-15
    norm = policy.ObsNormalizer(1)
    wts = _mean_net(d, "ck_general_")
    for t in range(len(d["rl_state"])):
        # current_k in the observation is the worker's stale local copy, NOT the live shared-memory parameters (:649, refreshed only at :896)
        obs = norm.push(d["rl_state"][t][None], d["rl_target"][t][None], d["rl_control"][t][None], cur_k[None])
        assert np.abs(obs - d["rl_obs"][t]).max() <= 1e-6, t
        mean = policy.mlp_forward(obs, wts)
        assert np.abs(mean - d["rl_mean"][t]).max() <= 1e-5
        if t % 8 == 0:
            k_new = policy.param_update(shm, d["rl_raw_action"][t])
            shm = policy.write_params(k_new, shm)
        assert rel(shm, d["rl_model_params"][t]) <= 1e-6, t
    assert np.abs(d["rl_obs"][:, -52 + 18:]).max() == 0.0           # the current_k slice is exactly zero after Welford normalisation


def test_ppo_update_matches_reference_rl_worker_run():
    """One real PPO update of the reference's worker (rollout 8, 2 epochs x 2 minibatches of 4): same buffer, same
    permutations -> same parameters."""
    import torch
    from oracle import ppo
    d = load("ref_policy.npz")
    lr, wd, clip_eps, vf, ent, gamma, lam, epochs, mb, npseed = d["ppo_hparams"]
    pol = ppo.Policy()
    pol.load_state_dict({k: torch.from_numpy(d["ppo_before_" + k.replace(".", "__")].copy()) for k in pol.state_dict()})
    opt = ppo.make_optimizer(pol, lr=float(lr), weight_decay=float(wd))
    obs = torch.from_numpy(d["ppo_buf_obs"].astype(np.float32))
    act = torch.from_numpy(d["ppo_buf_act"].astype(np.float32))
    logp = torch.from_numpy(d["ppo_buf_logp"].astype(np.float32))
    rewards, values, dones = d["ppo_buf_r_slot"].tolist(), d["ppo_buf_v_slot"].tolist(), d["ppo_buf_done"].tolist()
    with torch.no_grad():
        last_val = float(pol(obs[-1:])[2])          # :778-781: value of the last observation (obs of the filling step)
    adv = ppo.compute_gae(rewards, values, dones, last_val, float(gamma), float(lam))
    ret = torch.as_tensor(ppo.normalise_returns(np.array(adv) + np.array(values)), dtype=torch.float32)
    adv_t = ppo.normalise_advantages(adv)
    np.random.seed(int(npseed))
    for _ in range(int(epochs)):
        idxs = np.random.permutation(len(rewards))
        for s in range(0, len(rewards), int(mb)):
            ix = torch.from_numpy(idxs[s:s + int(mb)])
            ppo.minibatch_step(pol, opt, obs[ix], act[ix], logp[ix], adv_t[ix], ret[ix], clip_eps=float(clip_eps), vf_coef=float(vf), ent_coef=float(ent))
    worst = 0.0
    for k, v in pol.state_dict().items():
        a, b0 = d["ppo_after_" + k.replace(".", "__")], d["ppo_before_" + k.replace(".", "__")]
        worst = max(worst, np.abs(v.numpy() - a).max())
    assert worst <= 1e-6, worst
    assert np.all(d["ppo_adam_step"] == epochs * (len(rewards) // mb))
