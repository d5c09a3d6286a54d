"""Shared test helpers: host-emulation harness binding and oracle-side problem builders."""
import ctypes as C
import os

import numpy as np

import dart_b200
from oracle import problems

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NXF = {0: 6, 1: 4, 2: 8}

# parity bars of BASELINE.json north_star
TOL_U0 = 1e-4     # rad, first-move tilt command
TOL_J = 1e-6      # relative, optimal objective
TOL_RLS = 1e-6    # relative, RLS parameter trajectory


class HostEmu:
    """tests/hostemu/_build/libhostemu.so: the device solver source compiled for the host (1-lane tile)."""

    def __init__(self):
        self.lib = C.CDLL(os.path.join(ROOT, "tests", "hostemu", "_build", "libhostemu.so"))
        self.lib.hostemu_solve.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 9
        self.lib.hostemu_solve_dual.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 10
        self.lib.hostemu_ndual.argtypes = [C.c_void_p]

    def ndual(self, cfg):
        return self.lib.hostemu_ndual(C.byref(cfg))

    def solve(self, cfg, x0, ref, aux=None, warm=None, dual=None):
        x0 = np.ascontiguousarray(np.atleast_2d(x0), dtype=np.float64)
        B, N = x0.shape[0], cfg.N
        nw = (N + 1) * NXF[cfg.method] + 2 * N
        arrs = [x0] + [None if a is None else np.ascontiguousarray(np.atleast_2d(a), dtype=np.float64) for a in (ref, aux, warm)]
        w = np.zeros((B, nw)); u0 = np.zeros((B, 2)); J = np.zeros(B)
        st = np.zeros(B, np.int32); it = np.zeros(B, np.int32)
        p = lambda a: None if a is None else C.c_void_p(a.ctypes.data)
        if dual is None:
            rc = self.lib.hostemu_solve(C.byref(cfg), B, *[p(a) for a in arrs], p(w), p(u0), p(J), p(st), p(it))
        else:       # dual: float64 [B, ndual] array, read and updated in place (dart_set_dual_state semantics)
            assert dual.dtype == np.float64 and dual.flags.c_contiguous and dual.shape == (B, self.ndual(cfg))
            rc = self.lib.hostemu_solve_dual(C.byref(cfg), B, *[p(a) for a in arrs], p(w), p(u0), p(J), p(st), p(it), p(dual))
        assert rc == 0
        return dict(w=w, u0=u0, J=J, status=st, iters=it)


def pmpc_case(states_per_object=4, seed=1):
    c, aux = dart_b200.workloads.pmpc_inputs(states_per_object, seed)
    prob = problems.pmpc_problem(c["state"], c["target"], Qp=c["Qp"], Qv=c["Qv"], R=c["R"], mu=c["mu"])
    return c, aux, prob


def rmpc_case(B=32, seed=2, cold=False):
    """RMPC inputs mid-episode (dart_b200.workloads.rmpc_inputs) + the oracle's problem for them."""
    d = dart_b200.workloads.rmpc_inputs(B, seed, cold)
    return d, problems.rmpc_problem(d["x0"], d["u_prev"], d["theta"], d["ref"])


def lmpc_case(B=32, seed=3):
    d = dart_b200.workloads.lmpc_inputs(B, seed)
    return d, problems.lmpc_problem(d["x0"], d["u_prev"], d["pvec"], d["ref"])


def assert_parity(out, ref, what=""):
    """out: dict(u0, J, status) from the path under test; ref: oracle.ipm.solve result."""
    assert (ref["status"] == 0).all(), f"{what}: oracle did not converge"
    assert (out["status"] == 0).all(), f"{what}: statuses {np.bincount(out['status'])}"
    du0 = np.abs(out["u0"] - ref["U"][:, 0]).max()
    dJ = (np.abs(out["J"] - ref["J"]) / np.maximum(np.abs(ref["J"]), 1e-9)).max()
    assert du0 <= TOL_U0, f"{what}: |du0| = {du0:.3e} > {TOL_U0}"
    assert dJ <= TOL_J, f"{what}: rel dJ = {dJ:.3e} > {TOL_J}"
    return du0, dJ
