"""Generate tests/golden/*.npz from the oracle (run from the repo root: python tests/golden/make_golden.py).

PARITY UNPINNED by the reference: it ships no tests, fixtures or traces for this path and CasADi/IPOPT cannot be
installed here, so these vectors are the oracle's own converged KKT points (tol 1e-10, tighter than the 1e-8 the
solvers run at) on seeded inputs.  They pin the oracle against regressions and give the GPU tests a fixture that
does not need the oracle at run time.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import dart_b200                              # noqa: E402
from oracle import ipm, policy, problems, rls  # noqa: E402
from tests import helpers                      # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
TIGHT = ipm.Options(tol=1e-10)


def main():
    c, aux, p = helpers.pmpc_case(4)
    r = ipm.solve(p, opts=TIGHT)
    assert (r["status"] == 0).all()
    np.savez_compressed(os.path.join(OUT, "pmpc_config2_s4.npz"), state=c["state"], target=c["target"], aux=aux,
                        u0=r["U"][:, 0], J=r["J"], U=r["U"])
    c1 = dart_b200.workloads.pmpc_config1()
    r = ipm.solve(problems.pmpc_problem(c1["state"], c1["target"], Qp=400, Qv=2, R=0.2, mu=0.1), opts=TIGHT)
    np.savez_compressed(os.path.join(OUT, "pmpc_config1.npz"), state=c1["state"], target=c1["target"], u0=r["U"][:, 0], J=r["J"])

    d, p = helpers.rmpc_case(32)
    r = ipm.solve(p, opts=TIGHT)
    assert (r["status"] == 0).all()
    np.savez_compressed(os.path.join(OUT, "rmpc_b32.npz"), x0=d["x0"], ref=d["ref"], aux=d["aux"], u0=r["U"][:, 0], J=r["J"])

    d, p = helpers.lmpc_case(32)
    r = ipm.solve(p, opts=TIGHT)
    assert (r["status"] == 0).all()
    np.savez_compressed(os.path.join(OUT, "lmpc_b32.npz"), x0=d["x0"], ref=d["ref"], aux=d["aux"], u0=r["U"][:, 0], J=r["J"])

    rng = np.random.default_rng(11)
    B, T = 8, 200
    theta = np.zeros((B, 2, 7)); P = np.tile(np.eye(7) * 1e3, (B, 2, 1, 1))
    true = rng.standard_normal((B, 2, 7))
    phis, ys, traj = [], [], []
    for _ in range(T):
        x = 0.1 * rng.standard_normal((B, 4))
        phi = rls.regressor(x, 0.1)
        y = np.einsum("bep,bp->be", true, phi) + 1e-3 * rng.standard_normal((B, 2))
        theta, P = rls.rls_update_batch(theta, P, phi, y, 0.995)
        phis.append(phi); ys.append(y); traj.append(theta.copy())
    np.savez_compressed(os.path.join(OUT, "rls_traj.npz"), phi=np.array(phis), y=np.array(ys), theta=np.array(traj), P_final=P)

    weights = policy.orthogonal_policy_weights(seed=3)
    obs = np.random.default_rng(12).standard_normal((64, 520)).astype(np.float32)
    np.savez_compressed(os.path.join(OUT, "policy_mlp.npz"), obs=obs, mean=policy.mlp_forward(obs, weights, np.float64),
                        **{f"W{i}": W for i, (W, _) in enumerate(weights)}, **{f"b{i}": b for i, (_, b) in enumerate(weights)})
    print("golden vectors written to", OUT)


if __name__ == "__main__":
    main()
