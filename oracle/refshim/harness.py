"""Drivers that run the reference's controllers (loaded by ``loader``) deterministically.  TEST INFRASTRUCTURE ONLY.

* ``pmpc(...)`` / ``rmpc(...)``: construct ``PMPC`` / ``AdaptiveNPMPCSmooth`` on duck-typed MuJoCo objects.
* ``LmpcSolverWorker``: runs ``RLMPC._solver_worker`` (rlmpc2.py:228-533) unmodified in a thread, on real shared-memory
  segments, with a stepping gate in place of the ``state_ready`` event so that exactly one reference loop iteration runs
  per posted state (the reference's loop free-runs on a 10 ms timeout, which is not reproducible).
* ``RlWorker``: the same for ``RLMPC._rl_worker`` (rlmpc2.py:536-935), recording what the policy saw and produced.
* ``rlmpc_facade(...)``: an ``RLMPC`` object without its processes (``__init__`` bypassed), wired to a running
  ``LmpcSolverWorker`` so the reference's own ``RLMPC.solve`` (rlmpc2.py:986-1021) can be called.
"""
import threading
from multiprocessing import shared_memory

import numpy as np

from . import casadi as shim
from . import loader

LMPC_PACKET = {"Ts": 0.002, "nx": 8, "nu": 2, "N": 20, "Q": [200.0, 2.0, 200.0, 2.0, 0.0, 0.0, 0.0, 0.0],
               "Qt": [200.0, 2.0, 200.0, 2.0, 0.0, 0.0, 0.0, 0.0], "R": [0.1, 0.1, 1.0, 1.0],
               "u_bounds": (-0.4, 0.4), "params_len": 34, "g": 9.81}         # LMPC/src/run.py:118-126
LMPC_SHAPES = {"state": (8,), "state_next": (8,), "target": (8,), "w_opt": (8 * 21 + 2 * 20,), "loss": (1,),
               "control": (2,), "model_params": (34,), "state_deriv": (8,), "in_contact": (1,), "RLstatus": (1,)}   # rlmpc2.py:115-129


def set_pmpc_state(mpc, state):
    """What ``mpc_worker`` does before ``solve`` (mpc_3d.py:153-155)."""
    b = mpc.data.body(mpc.target_body)
    b.xpos[:] = [state[0], state[2], state[4]]
    b.cvel[3:6] = [state[1], state[3], state[5]]


def pmpc(Ts=0.002, **params):
    m = loader.load(loader.PMPC_FILE)
    return m.PMPC(loader.FakeModel(), loader.FakeData(), Ts, **params)


def rmpc(Ts=0.002, **params):
    m = loader.load(loader.RMPC_FILE)
    return m.AdaptiveNPMPCSmooth(loader.FakeModel(), loader.FakeData(), Ts, **params)


class Gate:
    """Event look-alike: ``wait`` blocks until the driver posts one step; ``is_waiting`` tells the driver the loop is
    parked at its ``wait`` again (i.e. the previous iteration has completed)."""

    def __init__(self):
        self._sem = threading.Semaphore(0)
        self._parked = threading.Event()

    def wait(self, timeout=None):
        self._parked.set()
        self._sem.acquire()
        return True

    def set(self):
        pass                        # RLMPC.solve sets state_ready itself; stepping is explicit (``post``)

    def post(self):
        self._parked.clear()
        self._sem.release()

    def until_parked(self, timeout=120.0):
        if not self._parked.wait(timeout):
            raise TimeoutError("reference worker did not return to its wait()")

    def clear(self):
        pass

    def is_set(self):
        return False


class _Shm:
    def __init__(self, shapes):
        self.shapes = dict(shapes)
        self.shms = {k: shared_memory.SharedMemory(create=True, size=max(8, int(np.prod(s)) * 8)) for k, s in shapes.items()}
        self.views = {k: np.ndarray(shapes[k], dtype=np.float64, buffer=self.shms[k].buf) for k in shapes}
        for v in self.views.values():
            v[:] = 0.0
        self.names = {k: s.name for k, s in self.shms.items()}

    def close(self):
        self.views = {}
        for s in self.shms.values():
            try:
                s.close()
                s.unlink()
            except Exception:
                pass


class LmpcSolverWorker:
    def __init__(self, packet=None, shm=None):
        self.mod = loader.load(loader.LMPC_FILE)
        self.packet = dict(LMPC_PACKET if packet is None else packet)
        self.own_shm = shm is None
        self.shm = _Shm(LMPC_SHAPES) if shm is None else shm
        self.views = self.shm.views
        self.events = {"state_ready": Gate(), "ctrl_ready": threading.Event(), "terminate": threading.Event(),
                       "reset": threading.Event(), "data_ready": threading.Event()}
        n0 = len(shim.CAPTURED)
        self.thread = threading.Thread(target=self.mod.RLMPC._solver_worker,
                                       args=(self.shm.names, self.events, self.packet, LMPC_SHAPES), daemon=True)
        self.thread.start()
        self.events["state_ready"].until_parked()
        self.solver = shim.CAPTURED[n0]            # the NLP the worker built (rlmpc2.py:478-491)

    def step(self, state, control, pvec, target):
        """One iteration of the reference's loop (:495-524): returns (w_opt, loss)."""
        v = self.views
        v["state"][:], v["control"][:], v["model_params"][:], v["target"][:] = state, control, pvec, target
        self.run_once()
        self.events["ctrl_ready"].clear()
        return v["w_opt"].copy(), v["loss"].copy()

    def run_once(self):
        """Let the worker solve whatever is in shared memory now; leaves ``ctrl_ready`` set, as the worker does."""
        g = self.events["state_ready"]
        g.post()
        g.until_parked()

    def close(self):
        self.events["terminate"].set()
        self.events["state_ready"].post()
        self.thread.join(10)
        if self.own_shm:
            self.shm.close()


def rlmpc_facade(worker, state_fn):
    """An ``RLMPC`` whose ``__init__`` (process + shm creation, rlmpc2.py:110-176) is bypassed; ``solve`` is the
    reference's.  ``state_fn() -> (8,)`` replaces the MuJoCo read of ``get_state``."""
    cls = worker.mod.RLMPC
    obj = object.__new__(cls)
    obj.params = dict(worker.packet)
    obj.views = worker.views
    obj.events = worker.events
    obj.last_control = np.array([0.0, 0.0])
    obj.get_state = state_fn
    return obj


class RlWorker:
    """``RLMPC._rl_worker`` in a thread; ``step()`` runs exactly one loop iteration and returns what was recorded."""

    def __init__(self, rl_packet, shm=None, torch_seed=0, numpy_seed=0):
        import torch
        self.mod = mod = loader.load(loader.LMPC_FILE)
        self.shm = _Shm(LMPC_SHAPES) if shm is None else shm
        self.own_shm = shm is None
        self.views = self.shm.views
        self.views["in_contact"][:] = 1.0
        self.events = {"state_ready": Gate(), "ctrl_ready": threading.Event(), "terminate": threading.Event(),
                       "reset": threading.Event(), "data_ready": threading.Event()}
        self.rec = rec = {"obs": [], "mean": [], "value": [], "raw_action": [], "policies": [], "optimizers": [], "buffer": []}

        # recorders around the reference's own classes (restored in close())
        self._orig_forward = mod.Policy.forward
        self._orig_init = mod.Policy.__init__
        self._orig_normal = mod.Normal
        self._orig_adam = mod.optim.Adam
        orig_forward, orig_init, OrigNormal, orig_adam = self._orig_forward, self._orig_init, self._orig_normal, self._orig_adam

        def forward(pol, obs):
            mean, std, value = orig_forward(pol, obs)
            if not torch.is_grad_enabled():
                rec["obs"].append(obs.detach().cpu().numpy().copy())
                rec["mean"].append(mean.detach().cpu().numpy().copy())
                rec["value"].append(value.detach().cpu().numpy().copy())
                rec["std"] = std.detach().cpu().numpy().copy()
            return mean, std, value

        def init(pol, *a, **k):
            orig_init(pol, *a, **k)
            rec["policies"].append(pol)

        class RecNormal(OrigNormal):
            def rsample(self_, *a, **k):
                s = OrigNormal.rsample(self_, *a, **k)
                rec["raw_action"].append(s.detach().cpu().numpy().copy())
                return s

        class _OptimNS:
            def __getattr__(self_, name):
                return getattr(self._orig_optim, name)

        def adam(*a, **k):
            o = orig_adam(*a, **k)
            rec["optimizers"].append(o)
            return o

        self._orig_add = mod.RolloutBuffer.add
        orig_add = self._orig_add

        def add(buf, o, a, logp, r, v, done):
            # the six values exactly as the worker passes them, by position (the worker passes value, reward in the
            # r, v slots -- rlmpc2.py:744 against :93); only the first buffer (``buf``, not ``glob_buf``) is recorded
            if not rec["buffer"] or rec["buffer"][0][0] is buf:
                rec["buffer"].append((buf, np.array(o), np.array(a), float(logp), float(r), float(v), float(done)))
            orig_add(buf, o, a, logp, r, v, done)

        mod.RolloutBuffer.add = add
        mod.Policy.forward = forward
        mod.Policy.__init__ = init
        mod.Normal = RecNormal
        self._orig_optim = mod.optim
        ns = type("optim_ns", (), {})()
        for nm in dir(mod.optim):
            if not nm.startswith("__"):
                setattr(ns, nm, getattr(mod.optim, nm))
        ns.Adam = adam
        mod.optim = ns
        torch.manual_seed(torch_seed)
        np.random.seed(numpy_seed)
        self.packet = dict(rl_packet)
        self.thread = threading.Thread(target=mod.RLMPC._rl_worker,
                                       args=(self.shm.names, self.events, self.packet, LMPC_SHAPES), daemon=True)
        self.thread.start()
        self.events["state_ready"].until_parked()

    def step(self, state, target, control):
        v = self.views
        v["state"][:], v["target"][:], v["control"][:] = state, target, control
        g = self.events["state_ready"]
        g.post()
        g.until_parked(600.0)
        return v["model_params"].copy()

    def close(self):
        self.events["terminate"].set()
        self.events["state_ready"].post()
        self.thread.join(10)
        mod = self.mod
        mod.Policy.forward, mod.Policy.__init__, mod.Normal, mod.optim = self._orig_forward, self._orig_init, self._orig_normal, self._orig_optim
        mod.RolloutBuffer.add = self._orig_add
        if self.own_shm:
            self.shm.close()
