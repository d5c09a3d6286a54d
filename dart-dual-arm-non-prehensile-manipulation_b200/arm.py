"""Low-level arm controller: the QP of ``PMPC/src/controller/arm.py`` for B arms at once.

The reference rebuilds a 7-variable NLP every 2 ms per arm from MuJoCo quantities and solves it with CasADi/IPOPT in a
worker process (``ARMCONTROL.solver_worker``, arm.py:265-457).  Here the host forms the QP data for the whole batch with
batched numpy linear algebra (``build_qp``, arm.py:337-405 vectorised) and ``dart_arm_qp_solve`` solves all of them in one
launch.  ``ArmQPBatch.solve(dyn)`` takes the dictionary ``ARMCONTROL.compute_dynamics`` returns (arm.py:186-200), with a
leading batch axis, and returns what the worker publishes: ``(torque [B,7], loss [B])`` and the accelerations.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import check

NV, NR = 7, 21


def default_params(dt=0.002):
    """arm.py:495-519 (R_params; L_params carry the same numbers)."""
    return dict(
        Wimp=np.diag([10.0, 10.0, 10.0, 1.0, 1.0, 1.0]), Wpos=np.eye(7) * 0.1, Wsmooth=np.eye(7) * 0.0,
        Qmin=np.array([-6.28319, -2.059, -6.28319, -0.19198, -6.28319, -1.69297, -6.28319]),
        Qmax=np.array([6.28319, 2.0944, 6.28319, 3.927, 6.28319, 3.14159, 6.28319]),
        Qdotmin=np.ones(7) * -20.0, Qdotmax=np.ones(7) * 20.0,
        taumin=np.array([-50.0, -50, -30, -30, -30, -20, -20]), taumax=np.array([50.0, 50, 30, 30, 30, 20, 20]),
        K=np.diag([1000.0, 1000.0, 1000.0, 50.0, 50.0, 50.0]) * 10, K_null=np.diag([1.0] * 7), dt=dt)


def _sqrtm_psd(A):
    """arm.py:363-366 for a stack of symmetric matrices: V sqrt(|w|) V'."""
    w, V = np.linalg.eigh(A)
    return (V * np.sqrt(np.abs(w))[:, None, :]) @ np.swapaxes(V, 1, 2)


def build_qp(dyn, params):
    """QP data of arm.py:337-405 for a batch.  Returns H [B,7,7], g [B,7], c0 [B], C [B,21,7], lo [B,21], hi [B,21]
    with cost = 0.5 x'Hx + g'x + c0 (the reference's loss) and lo <= Cx <= hi."""
    f = lambda k: np.asarray(dyn[k], dtype=np.float64)
    q, qd, qdd_prev = f("q"), f("qd"), f("qdd_prev")
    jac, jacDot, M, h, Mx_inv = f("jac"), f("jacDot"), f("M"), f("h"), f("Mx_inv")
    B = q.shape[0]
    K, K_null, dt = np.asarray(params["K"], float), np.asarray(params["K_null"], float), float(params["dt"])
    twist = np.concatenate([f("mocap_pos") - f("ee_pos"), f("rotvec")], axis=1)
    Minv = np.linalg.pinv(M, rcond=1e-6)                                          # :346-349
    det = np.linalg.det(Mx_inv)                                                    # :351-357
    Mx = np.empty_like(Mx_inv)
    reg = np.abs(det) > 1e-8
    if reg.any():
        Mx[reg] = np.linalg.inv(Mx_inv[reg])
    if (~reg).any():
        Mx[~reg] = np.linalg.pinv(Mx_inv[~reg], rcond=1e-3)
    mv = lambda A, x: np.einsum('bij,bj->bi', A, x)
    mu = mv(Mx, mv(jac, mv(Minv, h)) + mv(jacDot, qd))                              # :360
    sMx, sK = _sqrtm_psd(Mx), np.sqrt(K)
    D = sMx @ sK + sK @ sMx                                                        # :368-370
    F = -mv(D, mv(jac, qd)) + twist @ K.T + mu                                     # :384
    e0 = mv(jacDot, qd) - mv(Mx_inv, F)                                            # Eimp = jac qdd + e0 (:385)
    beta = 2.0 * np.sqrt(np.diag(K_null)) * (-qd) + (-q) @ K_null.T                # :387
    Wimp, Wpos = np.asarray(params["Wimp"], float), np.asarray(params["Wpos"], float)
    Wsm = np.asarray(params["Wsmooth"], float) / dt ** 2
    jT = np.swapaxes(jac, 1, 2)
    H = 2.0 * (jT @ Wimp @ jac + Wpos + Wsm)
    H = 0.5 * (H + np.swapaxes(H, 1, 2))
    g = 2.0 * (mv(jT, e0 @ Wimp.T) - beta @ Wpos.T - qdd_prev @ Wsm.T)
    c0 = np.einsum('bi,ij,bj->b', e0, Wimp, e0) + np.einsum('bi,ij,bj->b', beta, Wpos, beta) + \
        np.einsum('bi,ij,bj->b', qdd_prev, Wsm, qdd_prev)
    I = np.broadcast_to(np.eye(7), (B, 7, 7))
    Cm = np.concatenate([0.5 * dt ** 2 * I, dt * I, M], axis=1)                    # :399-402
    off = np.concatenate([qd * dt + q, qd, h], axis=1)
    lo = np.concatenate([params["Qmin"], params["Qdotmin"], params["taumin"]])[None, :] - off
    hi = np.concatenate([params["Qmax"], params["Qdotmax"], params["taumax"]])[None, :] - off
    return H, g, c0, np.ascontiguousarray(Cm), lo, hi


def solve_qp_device(H, g, Cm, lo, hi, x0=None, tol=1e-8, max_iter=100, out=None):
    """CUDA tensors in, CUDA tensors out; one launch on the current stream of the tensors' device."""
    import torch
    B = H.shape[0]
    dev = H.device
    if out is None:
        out = dict(x=torch.empty((B, NV), dtype=torch.float64, device=dev), obj=torch.empty((B,), dtype=torch.float64, device=dev),
                   status=torch.empty((B,), dtype=torch.int32, device=dev), iters=torch.empty((B,), dtype=torch.int32, device=dev))
    for t in (H, g, Cm, lo, hi) + ((x0,) if x0 is not None else ()):
        if not (t.is_cuda and t.dtype == torch.float64 and t.is_contiguous()):
            raise ValueError("dart_arm_qp_solve takes contiguous float64 CUDA tensors")
    p = lambda t: None if t is None else C.c_void_p(t.data_ptr())
    with torch.cuda.device(dev):
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        check(_lib.lib().dart_arm_qp_solve(B, p(H), p(g), p(Cm), p(lo), p(hi), p(x0), p(out["x"]), p(out["obj"]),
                                            p(out["status"]), p(out["iters"]), float(tol), int(max_iter), stream),
              "dart_arm_qp_solve")
    return out


_DYN_KEYS = ("q", "qd", "qdd_prev", "mocap_pos", "ee_pos", "rotvec", "jac", "jacDot", "M", "h", "Mx_inv")
_DYN_SHAPES = {"q": (7,), "qd": (7,), "qdd_prev": (7,), "mocap_pos": (3,), "ee_pos": (3,), "rotvec": (3,), "jac": (6, 7),
               "jacDot": (6, 7), "M": (7, 7), "h": (7,), "Mx_inv": (6, 6)}


def build_qp_device(dyn, params, out=None):
    """``build_qp`` on the GPU (dart_arm_qp_build): ``dyn`` holds float64 CUDA tensors with a leading batch axis."""
    import torch
    B = dyn["q"].shape[0]
    dev = dyn["q"].device
    for k in _DYN_KEYS:
        t = dyn[k]
        if not (t.is_cuda and t.dtype == torch.float64 and t.is_contiguous() and tuple(t.shape) == (B,) + _DYN_SHAPES[k]):
            raise ValueError(f"dyn['{k}'] must be a contiguous float64 CUDA tensor of shape {(B,) + _DYN_SHAPES[k]}")
    if out is None:
        e = lambda *s: torch.empty((B,) + s, dtype=torch.float64, device=dev)
        out = dict(H=e(NV, NV), g=e(NV), c0=e(), C=e(NR, NV), lo=e(NR), hi=e(NR))
    hp = lambda a, n: np.ascontiguousarray(np.asarray(a, dtype=np.float64).reshape(n))
    host = [hp(params["Wimp"], 36), hp(params["Wpos"], 49), hp(params["Wsmooth"], 49), hp(params["K"], 36), hp(params["K_null"], 49),
            hp(np.concatenate([params["Qmin"], params["Qdotmin"], params["taumin"]]), 21),
            hp(np.concatenate([params["Qmax"], params["Qdotmax"], params["taumax"]]), 21)]
    p = lambda t: C.c_void_p(t.data_ptr())
    with torch.cuda.device(dev):
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        check(_lib.lib().dart_arm_qp_build(B, *[C.c_void_p(a.ctypes.data) for a in host], float(params["dt"]),
                                            *[p(dyn[k]) for k in _DYN_KEYS], p(out["H"]), p(out["g"]), p(out["c0"]),
                                            p(out["C"]), p(out["lo"]), p(out["hi"]), stream), "dart_arm_qp_build")
    return out


class ArmQPBatch:
    """B arm controllers: ``solve(dyn)`` = one cycle of the reference's solver worker for each of them (primal warm start
    from the previous accelerations, arm.py:412-418; IPOPT's dual warm start has no counterpart -- cold duals)."""

    def __init__(self, params=None, device=0, tol=1e-8, max_iter=100, warm_start=True):
        import torch
        self.torch = torch
        self.params = dict(default_params() if params is None else params)
        self.dev = torch.device("cuda", device)
        self.tol, self.max_iter, self.warm_start = tol, max_iter, warm_start
        self.prev = None
        self.launches = 0

    def solve_device(self, dyn):
        """``dyn``: CUDA tensors.  Two launches (QP build, QP solve) + one batched matrix-vector product for the torque;
        returns device tensors (tau [B,7], loss [B], qdd [B,7]); ``self.status`` / ``self.iters`` stay on the device."""
        torch = self.torch
        qp = build_qp_device(dyn, self.params)
        B = qp["H"].shape[0]
        x0 = self.prev if (self.warm_start and self.prev is not None and self.prev.shape[0] == B) else None
        out = solve_qp_device(qp["H"], qp["g"], qp["C"], qp["lo"], qp["hi"], x0=x0, tol=self.tol, max_iter=self.max_iter)
        self.launches += 2
        self.prev = out["x"]
        tau = torch.bmm(dyn["M"], out["x"][:, :, None])[:, :, 0] + dyn["h"]         # arm.py:425
        self.status, self.iters = out["status"], out["iters"]
        return tau, out["obj"] + qp["c0"], out["x"]

    def _staging(self, B):
        """One pinned host buffer / one device buffer holding all eleven inputs back to back (contiguous per key), and
        one pair for the outputs [tau | loss | qdd | status | iters]: a control cycle costs one H2D and one D2H copy."""
        torch = self.torch
        if getattr(self, "_stage_B", None) != B:
            sizes = [int(np.prod(_DYN_SHAPES[k])) for k in _DYN_KEYS]
            self._offs = np.concatenate([[0], np.cumsum([B * sz for sz in sizes])])
            self._pin_in = torch.empty(int(self._offs[-1]), dtype=torch.float64).pin_memory()
            self._dev_in = torch.empty(int(self._offs[-1]), dtype=torch.float64, device=self.dev)
            self._np_in = self._pin_in.numpy()
            self._dev_out = torch.empty((B, 17), dtype=torch.float64, device=self.dev)
            self._pin_out = torch.empty((B, 17), dtype=torch.float64).pin_memory()
            self._stage_B = B
        return self._offs

    def solve(self, dyn):
        """``dyn``: numpy arrays with a leading batch axis (what compute_dynamics returns, stacked).  Host in, host out."""
        torch = self.torch
        B = np.asarray(dyn["q"]).shape[0]
        offs = self._staging(B)
        d = {}
        for i, k in enumerate(_DYN_KEYS):
            self._np_in[offs[i]:offs[i + 1]] = np.asarray(dyn[k], dtype=np.float64).reshape(-1)
        self._dev_in.copy_(self._pin_in, non_blocking=True)
        for i, k in enumerate(_DYN_KEYS):
            d[k] = self._dev_in[offs[i]:offs[i + 1]].view((B,) + _DYN_SHAPES[k])
        tau, loss, x = self.solve_device(d)
        o = self._dev_out
        o[:, 0:7] = tau; o[:, 7] = loss; o[:, 8:15] = x; o[:, 15] = self.status; o[:, 16] = self.iters
        self._pin_out.copy_(o, non_blocking=True)
        torch.cuda.current_stream(self.dev).synchronize()
        h = self._pin_out.numpy()
        self.status, self.iters = h[:, 15].astype(np.int32), h[:, 16].astype(np.int32)
        return h[:, 0:7].copy(), h[:, 7].copy(), h[:, 8:15].copy()


class ARMCONTROL:
    """Facade with the reference's constructor and ``compute_torque`` (arm.py:17-110, 201-232), synchronous: the torque of
    THIS cycle's QP is returned (the reference returns whatever its worker finished within 5 ms).  ``compute_dynamics``
    needs the ``mujoco`` module exactly as the reference does; ``compute_torque_from(dynamics)`` takes the dictionary
    directly."""

    def __init__(self, model, data, params, device=0):
        self.model, self.data, self.params = model, data, dict(params)
        self.nq = len(params.get("joint_names", range(7)))
        self._batch = ArmQPBatch(self.params, device=device)
        self.views = {"qdd_prev": np.zeros(self.nq), "torque_out": np.zeros(self.nq), "loss_out": np.zeros(1)}

    def compute_dynamics(self):
        raise NotImplementedError("compute_dynamics needs MuJoCo (mj_jacBody, mj_fullM, mj_solveM, mj_jacDot; arm.py:111-200); "
                                  "pass its dictionary to compute_torque_from()")

    def compute_torque_from(self, dynamics):
        dyn = {k: np.asarray(v, dtype=np.float64)[None] for k, v in dynamics.items() if k in (
            "q", "qd", "qdd_prev", "jac", "jacDot", "M", "h", "Mx_inv", "ee_pos", "mocap_pos", "rotvec")}
        tau, loss, x = self._batch.solve(dyn)
        if self._batch.status[0] in (_lib.STATUS_CONVERGED, _lib.STATUS_ACCEPTABLE):
            self.views["qdd_prev"][:] = x[0]
            self.views["torque_out"][:] = tau[0]
            self.views["loss_out"][0] = loss[0]
        else:                                       # arm.py:433-437: keep the previous accelerations, zero torque, loss -3
            self.views["torque_out"][:] = 0.0
            self.views["loss_out"][0] = -3.0
        return self.views["torque_out"].copy(), float(self.views["loss_out"][0])

    def compute_torque(self):
        return self.compute_torque_from(self.compute_dynamics())

    def close(self):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
