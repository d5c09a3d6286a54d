"""Device-resident PMPC closed loop on the surrogate plant vs the oracle's loop, plus long-episode properties."""
import numpy as np
import pytest

import dart_b200
from oracle import closed_loop
from tests import helpers

pytestmark = pytest.mark.gpu


def test_short_episode_matches_oracle_loop(built):
    c, aux, _ = helpers.pmpc_case(1)
    B, T = 18, 15
    rng = np.random.default_rng(0)
    coul = rng.uniform(0.0, 0.05, B)
    ep = dart_b200.PMPCEpisodes(c["state"], c["target"], aux, coulomb=coul, device=0)
    m = ep.run(T)
    o = closed_loop.pmpc_episode(c["state"], c["target"], aux, T, coulomb=coul)
    # commands agree to <= 1e-4 rad per step: x/y velocities then differ by <= Ts*g*1e-4 = 2e-6 per step, the z rows
    # (vz_new = -g |theta|^2) by <= 2*g*0.6*1e-4 = 1.2e-3
    d = np.abs(ep.state.cpu().numpy() - o["state"])
    assert d[:, :4].max() < 5e-5 and d[:, 4:].max() < 5e-3, (d[:, :4].max(), d[:, 4:].max())
    assert np.abs(ep.u0.cpu().numpy() - o["u"][-1]).max() <= helpers.TOL_U0
    assert np.abs(m["control_effort"] - o["effort"]).max() < 1e-6
    assert m["not_converged_solves"] == 0


def test_config1_episode_settles(built):
    """BASELINE config 1 (cube, mu = 0.10, main.py weights) on the model-as-plant: 10 cm move settles within 1 cm."""
    c1 = dart_b200.workloads.pmpc_config1()
    aux = np.stack([c1["Qp"], c1["Qv"], c1["R"], c1["mu"]], axis=1)
    ep = dart_b200.PMPCEpisodes(c1["state"], c1["target"], aux, device=0)
    m = ep.run(2500, trace_every=50)
    assert m["not_converged_solves"] == 0
    assert m["converged"][0] and m["convergence_time"][0] < 5.0
    assert m["steady_state_error"][0] < 0.01
    tr = m["trace"]
    assert tr.shape[1] == 11 and np.all(np.abs(tr[:, 7:9]) <= 0.6 + 1e-12)
    err = np.hypot(tr[:, 1] - 0.1, tr[:, 3] - 0.05)
    assert err[-1] < err[0] * 0.1


def test_graph_replay_equals_eager(built):
    """One closed-loop step captured in a CUDA graph and replayed gives bit-identical episodes."""
    c, aux, _ = helpers.pmpc_case(2)
    a = dart_b200.PMPCEpisodes(c["state"], c["target"], aux, device=0)
    b = dart_b200.PMPCEpisodes(c["state"], c["target"], aux, device=0)
    ma = a.run(60)
    mb = b.run(60, graph=True)
    assert np.array_equal(a.state.cpu().numpy(), b.state.cpu().numpy())
    assert np.array_equal(ma["control_effort"], mb["control_effort"]) and np.array_equal(ma["convergence_time"], mb["convergence_time"])
    assert ma["mean_iters"] == mb["mean_iters"] and mb["not_converged_solves"] == 0
    assert (b.nsteps.cpu().numpy() == 60).all()


def test_reference_main_loop_example(built):
    """examples/pmpc_main_surrogate.py: the reference's main.py loop with the drop-in class settles the object."""
    import importlib.util, os
    spec = importlib.util.spec_from_file_location("ex", os.path.join(helpers.ROOT, "examples", "pmpc_main_surrogate.py"))
    ex = importlib.util.module_from_spec(spec); spec.loader.exec_module(ex)
    log = ex.run([0.1, 0.0, 0.05, 0.0, 0.0, 0.0], friction=0.1, steps=900, verbose=False)
    assert (log[:, 6] == 0).all()                   # every solve converged
    assert log[-1, 1] < 0.01 and log[0, 1] > 0.1    # 11 cm -> under 1 cm
    assert np.abs(log[:, 2:4]).max() <= 0.6 + 1e-12


def test_persistent_episode_is_bit_identical_to_stepwise(built):
    """dart_pmpc_episode (all steps in one launch) against solve + plant launched step by step."""
    c, aux = dart_b200.workloads.pmpc_inputs(4)
    rng = np.random.default_rng(3)
    cou = rng.uniform(0, 0.02, aux.shape[0])
    a = dart_b200.PMPCEpisodes(c["state"], c["target"], aux, coulomb=cou, device=0)
    b = dart_b200.PMPCEpisodes(c["state"], c["target"], aux, coulomb=cou, device=0)
    ma = a.run(150)
    mb = b.run(150, persistent=True)
    assert np.array_equal(a.state.cpu().numpy(), b.state.cpu().numpy())
    for k in ("steady_state_error", "convergence_time", "control_effort"):
        assert np.array_equal(ma[k], mb[k]), k
    assert ma["mean_iters"] == mb["mean_iters"] and ma["not_converged_solves"] == mb["not_converged_solves"] == 0
    assert np.array_equal(a.u0.cpu().numpy(), b.u0.cpu().numpy())
    # a second call continues the same episode
    a.run(20); b.run(20, persistent=True)
    assert np.array_equal(a.state.cpu().numpy(), b.state.cpu().numpy()) and a.step_index == b.step_index == 170
