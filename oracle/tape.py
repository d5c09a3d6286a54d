"""Expression tapes: the reference's own CasADi graphs, captured and replayed with numpy.  TEST INFRASTRUCTURE ONLY.

The reference builds its NLPs by running Python over ``casadi.SX`` objects (mpc_3d.py:28-85,
np_mpc_adaptive_with_linear_regressor.py:65-168, rlmpc2.py:236-491).  CasADi itself is not installable here, but that
Python is: ``oracle/refshim`` provides a stand-in ``casadi`` module whose ``SX`` records every scalar operation the
reference performs into a tape (a straight-line program over + - * / sin tanh exp fabs ...).  A tape is therefore the
reference's expression graph, produced by executing the reference's source, and it travels as a small ``.npz`` fixture
(``tests/golden/ref_*_tape.npz``) to machines that do not have ``/root/reference`` -- the GPU box evaluates the
*reference's* objective, constraints and their derivatives at the CUDA solver's solutions by replaying it.

This file is the replayer: numpy only, batched over leading axes, complex-safe (so first derivatives come from the
complex-step method, exact to rounding), with a reverse sweep for gradients / vector-Jacobian products.
"""
import numpy as np

# opcodes (binary ops use a, b; unary use a)
CONST, INPUT, ADD, SUB, MUL, DIV, NEG, SIN, COS, TANH, EXP, FABS, SQ, SQRT, LOG = range(15)
OPNAMES = ["const", "input", "add", "sub", "mul", "div", "neg", "sin", "cos", "tanh", "exp", "fabs", "sq", "sqrt", "log"]
_BINARY = (ADD, SUB, MUL, DIV)


def _fabs(x):
    if np.iscomplexobj(x):
        return np.where(x.real < 0, -x, x)
    return np.abs(x)


class Tape:
    """Straight-line program.  ``op/a/b`` int arrays over nodes (topologically ordered), ``cval`` the constants,
    ``inputs``/``outputs``: name -> int array of node ids (inputs are INPUT nodes)."""

    def __init__(self, op, a, b, cval, inputs, outputs, meta=None):
        self.op = np.asarray(op, np.int16)
        self.a = np.asarray(a, np.int32)
        self.b = np.asarray(b, np.int32)
        self.cval = np.asarray(cval, np.float64)
        self.inputs = {k: np.asarray(v, np.int32) for k, v in inputs.items()}
        self.outputs = {k: np.asarray(v, np.int32) for k, v in outputs.items()}
        self.meta = dict(meta or {})
        self._sched = None

    # ---- persistence -------------------------------------------------------------------------------------------
    def save(self, path):
        d = {"op": self.op, "a": self.a, "b": self.b, "cval": self.cval}
        for k, v in self.inputs.items():
            d["in__" + k] = v
        for k, v in self.outputs.items():
            d["out__" + k] = v
        for k, v in self.meta.items():
            d["meta__" + k] = np.asarray(v)
        np.savez_compressed(path, **d)

    @classmethod
    def load(cls, path):
        z = np.load(path, allow_pickle=False)
        ins = {k[4:]: z[k] for k in z.files if k.startswith("in__")}
        outs = {k[5:]: z[k] for k in z.files if k.startswith("out__")}
        meta = {k[6:]: z[k] for k in z.files if k.startswith("meta__")}
        return cls(z["op"], z["a"], z["b"], z["cval"], ins, outs, meta)

    # ---- scheduling: nodes grouped by (level, op) so one numpy call handles a whole group ------------------------
    def _schedule(self):
        if self._sched is not None:
            return self._sched
        n = len(self.op)
        level = np.zeros(n, np.int32)
        op, a, b = self.op, self.a, self.b
        for i in range(n):
            o = op[i]
            if o <= INPUT:
                continue
            l = level[a[i]]
            if o in _BINARY and level[b[i]] > l:
                l = level[b[i]]
            level[i] = l + 1
        groups = []
        idx = np.nonzero(op > INPUT)[0]
        if len(idx):
            key = level[idx].astype(np.int64) * 32 + op[idx]
            order = np.argsort(key, kind="stable")
            idx, key = idx[order], key[order]
            cuts = np.nonzero(np.diff(key))[0] + 1
            for seg in np.split(idx, cuts):
                groups.append((int(op[seg[0]]), seg, a[seg], b[seg]))
        self._sched = groups
        return groups

    # ---- forward ------------------------------------------------------------------------------------------------
    def _forward(self, feeds):
        """-> vals [n_nodes, *batch].  feeds: name -> array [*batch, len(inputs[name])] (broadcast over batch)."""
        arrs = {k: np.asarray(v) for k, v in feeds.items()}
        for k in self.inputs:
            if k not in arrs:
                raise KeyError(f"tape input {k!r} missing")
        bshape = np.broadcast_shapes(*[v.shape[:-1] for v in arrs.values()])
        dtype = np.result_type(np.float64, *[v.dtype for v in arrs.values()])
        vals = np.empty((len(self.op),) + bshape, dtype)
        cmask = self.op == CONST
        vals[cmask] = self.cval[cmask].reshape((-1,) + (1,) * len(bshape))
        for k, ids in self.inputs.items():
            v = np.broadcast_to(arrs[k], bshape + (len(ids),))
            vals[ids] = np.moveaxis(v, -1, 0)
        for o, seg, ia, ib in self._schedule():
            x = vals[ia]
            if o == ADD:
                r = x + vals[ib]
            elif o == SUB:
                r = x - vals[ib]
            elif o == MUL:
                r = x * vals[ib]
            elif o == DIV:
                r = x / vals[ib]
            elif o == NEG:
                r = -x
            elif o == SIN:
                r = np.sin(x)
            elif o == COS:
                r = np.cos(x)
            elif o == TANH:
                r = np.tanh(x)
            elif o == EXP:
                r = np.exp(x)
            elif o == FABS:
                r = _fabs(x)
            elif o == SQ:
                r = x * x
            elif o == SQRT:
                r = np.sqrt(x)
            elif o == LOG:
                r = np.log(x)
            else:
                raise ValueError(f"bad opcode {o}")
            vals[seg] = r
        return vals

    def eval(self, **feeds):
        """-> {output name: array [*batch, n_out]}."""
        vals = self._forward(feeds)
        return {k: np.moveaxis(vals[ids], 0, -1) for k, ids in self.outputs.items()}

    # ---- reverse sweep --------------------------------------------------------------------------------------------
    def vjp(self, seeds, wrt, **feeds):
        """Sum over outputs of seed . d(output)/d(wrt).  seeds: {output name: [*batch, n_out]} -> [*batch, n_wrt].
        Complex feeds are allowed (the result is then the analytic continuation: complex-step it for second derivatives)."""
        vals = self._forward(feeds)
        adj = np.zeros_like(vals)
        for k, s in seeds.items():
            s = np.broadcast_to(np.asarray(s), vals.shape[1:] + (len(self.outputs[k]),))
            np.add.at(adj, self.outputs[k], np.moveaxis(s, -1, 0))
        for o, seg, ia, ib in reversed(self._schedule()):
            g = adj[seg]
            if o == ADD:
                np.add.at(adj, ia, g)
                np.add.at(adj, ib, g)
            elif o == SUB:
                np.add.at(adj, ia, g)
                np.add.at(adj, ib, -g)
            elif o == MUL:
                np.add.at(adj, ia, g * vals[ib])
                np.add.at(adj, ib, g * vals[ia])
            elif o == DIV:
                q = g / vals[ib]
                np.add.at(adj, ia, q)
                np.add.at(adj, ib, -q * vals[seg])
            elif o == NEG:
                np.add.at(adj, ia, -g)
            elif o == SIN:
                np.add.at(adj, ia, g * np.cos(vals[ia]))
            elif o == COS:
                np.add.at(adj, ia, -g * np.sin(vals[ia]))
            elif o == TANH:
                np.add.at(adj, ia, g * (1.0 - vals[seg] * vals[seg]))
            elif o == EXP:
                np.add.at(adj, ia, g * vals[seg])
            elif o == FABS:
                np.add.at(adj, ia, g * np.where(vals[ia].real < 0, -1.0, 1.0))
            elif o == SQ:
                np.add.at(adj, ia, 2.0 * g * vals[ia])
            elif o == SQRT:
                np.add.at(adj, ia, 0.5 * g / vals[seg])
            elif o == LOG:
                np.add.at(adj, ia, g / vals[ia])
        return np.moveaxis(adj[self.inputs[wrt]], 0, -1)

    def grad(self, out, wrt, **feeds):
        """Gradient of a scalar output."""
        return self.vjp({out: np.ones(1)}, wrt, **feeds)

    def jac(self, out, wrt, **feeds):
        """Dense Jacobian [*batch, n_out, n_wrt] by the complex-step method (one batched forward pass)."""
        h = 1e-30
        x = np.asarray(feeds[wrt], np.float64)
        n = x.shape[-1]
        f2 = {k: np.asarray(v)[..., None, :] for k, v in feeds.items()}
        f2[wrt] = x[..., None, :] + 1j * h * np.eye(n)
        vals = self._forward(f2)
        J = np.moveaxis(vals[self.outputs[out]], 0, -1).imag / h          # [*batch, n_wrt, n_out]
        return np.swapaxes(J, -1, -2)

    def hess_lagrangian(self, seeds, wrt, **feeds):
        """d/d(wrt) of vjp(seeds, wrt): the Hessian of sum_k seed_k . output_k, [*batch, n, n] (complex step over the reverse sweep)."""
        h = 1e-30
        x = np.asarray(feeds[wrt], np.float64)
        n = x.shape[-1]
        f2 = {k: np.asarray(v)[..., None, :] for k, v in feeds.items()}
        f2[wrt] = x[..., None, :] + 1j * h * np.eye(n)
        s2 = {k: np.asarray(v)[..., None, :] for k, v in seeds.items()}
        G = self.vjp(s2, wrt, **f2)                                        # [*batch, n(dir), n]
        H = G.imag / h
        return 0.5 * (H + np.swapaxes(H, -1, -2))

    def __repr__(self):
        return (f"Tape({len(self.op)} nodes, in={ {k: len(v) for k, v in self.inputs.items()} }, "
                f"out={ {k: len(v) for k, v in self.outputs.items()} })")
