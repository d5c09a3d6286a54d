"""Stand-in for the ``casadi`` module so the reference's own Python can be executed here.  TEST INFRASTRUCTURE ONLY.

CasADi (``casadi>=3.5.0``, PMPC/requirements.txt:10) is not installed and cannot be installed (no network).  The
reference files on the hot path use a small part of its API (SX.sym, Function, vertcat/vertsplit/reshape, sumsqr, dot,
mtimes, diag, inv, fabs/sin/tanh/exp, inf, nlpsol).  This module implements that part:

* ``SX`` is a dense matrix of scalar graph nodes; every arithmetic operation the reference performs is recorded in a
  global, hash-consed straight-line program.  ``Function`` / ``nlpsol`` turn sub-graphs into ``oracle.tape.Tape``
  objects, i.e. the reference's expression graphs, replayable with numpy.
* ``nlpsol(name, 'ipopt', nlp, opts)`` *captures* ``{x, f, g, p}`` (every instance is appended to ``CAPTURED``) and
  returns a callable with CasADi's calling convention (x0, p, lbx, ubx, lbg, ubg -> {'x','f','g','lam_g','lam_x'}).
  The solve itself is NOT IPOPT: it is ``oracle.refshim.nlp_solve`` (a dense primal-dual interior-point method on the
  captured graphs, exact derivatives).  These NLPs are small and locally convex, so their KKT point does not depend on
  the solver; what the fixtures pin is the reference's *problem* (model, cost, constraints, bounds, parameter layout,
  warm-start and calling conventions), all of it produced by running the reference's source unmodified.

Loaded under the name ``casadi`` by ``oracle.refshim.loader`` only; nothing in the product imports it.
"""
import math

import numpy as np

from .. import tape as _t

inf = float("inf")
pi = math.pi


# ---------------------------------------------------------------------------------------------------------------------
# global graph
# ---------------------------------------------------------------------------------------------------------------------
class _Graph:
    def __init__(self):
        self.op, self.a, self.b, self.cval = [], [], [], []
        self.names = {}
        self._cons = {}
        self._const = {}

    def const(self, v):
        v = float(v)
        key = v.hex() if v == v else "nan"
        i = self._const.get(key)
        if i is None:
            i = self._push(_t.CONST, -1, -1, v)
            self._const[key] = i
        return i

    def _push(self, op, a, b, c=0.0):
        self.op.append(op)
        self.a.append(a)
        self.b.append(b)
        self.cval.append(c)
        return len(self.op) - 1

    def sym(self, name):
        i = self._push(_t.INPUT, -1, -1)
        self.names[i] = name
        return i

    def node(self, op, a, b=-1):
        key = (op, a, b)
        i = self._cons.get(key)
        if i is None:
            i = self._push(op, a, b)
            self._cons[key] = i
        return i

    def is_const(self, i):
        return self.op[i] == _t.CONST


G = _Graph()


class S:
    """Scalar graph node with Python operator overloading."""
    __slots__ = ("i",)
    __array_ufunc__ = None

    def __init__(self, i):
        self.i = i

    @staticmethod
    def of(v):
        if isinstance(v, S):
            return v
        if isinstance(v, SX):
            if v.m.size != 1:
                raise ValueError("matrix used where a scalar is expected")
            return v.m.flat[0]
        return S(G.const(float(v)))

    def cv(self):
        return G.cval[self.i] if G.op[self.i] == _t.CONST else None

    # binary ops with constant folding and the identities CasADi applies too (x+0, x*1, x*0, 0/x)
    def _bin(self, o, op, swap=False):
        o = S.of(o)
        x, y = (o, self) if swap else (self, o)
        cx, cy = x.cv(), y.cv()
        if cx is not None and cy is not None:
            if op == _t.ADD:
                return S(G.const(cx + cy))
            if op == _t.SUB:
                return S(G.const(cx - cy))
            if op == _t.MUL:
                return S(G.const(cx * cy))
            if op == _t.DIV:
                return S(G.const(cx / cy))
        if op == _t.ADD:
            if cx == 0.0:
                return y
            if cy == 0.0:
                return x
        elif op == _t.SUB:
            if cy == 0.0:
                return x
            if cx == 0.0:
                return -y
        elif op == _t.MUL:
            if cx == 0.0 or cy == 0.0:
                return S(G.const(0.0))
            if cx == 1.0:
                return y
            if cy == 1.0:
                return x
        elif op == _t.DIV:
            if cx == 0.0:
                return S(G.const(0.0))
            if cy == 1.0:
                return x
        return S(G.node(op, x.i, y.i))

    def __add__(self, o): return self._bin(o, _t.ADD)
    def __radd__(self, o): return self._bin(o, _t.ADD, True)
    def __sub__(self, o): return self._bin(o, _t.SUB)
    def __rsub__(self, o): return self._bin(o, _t.SUB, True)
    def __mul__(self, o): return self._bin(o, _t.MUL)
    def __rmul__(self, o): return self._bin(o, _t.MUL, True)
    def __truediv__(self, o): return self._bin(o, _t.DIV)
    def __rtruediv__(self, o): return self._bin(o, _t.DIV, True)

    def __neg__(self):
        c = self.cv()
        if c is not None:
            return S(G.const(-c))
        return S(G.node(_t.NEG, self.i))

    def __pos__(self):
        return self

    def __pow__(self, e):
        if isinstance(e, (S, SX)):
            ce = S.of(e).cv()
            if ce is None:
                raise NotImplementedError("symbolic exponent")
            e = ce
        e = float(e)
        if e == 2.0:
            return self.un(_t.SQ)
        if e == 1.0:
            return self
        if e == 0.5:
            return self.un(_t.SQRT)
        if e == int(e) and 0 < e <= 8:
            r = self
            for _ in range(int(e) - 1):
                r = r * self
            return r
        raise NotImplementedError(f"power {e}")

    def un(self, op):
        c = self.cv()
        if c is not None:
            f = {_t.SIN: math.sin, _t.COS: math.cos, _t.TANH: math.tanh, _t.EXP: math.exp, _t.FABS: abs,
                 _t.SQ: lambda v: v * v, _t.SQRT: math.sqrt, _t.LOG: math.log}[op]
            return S(G.const(f(c)))
        return S(G.node(op, self.i))


# ---------------------------------------------------------------------------------------------------------------------
# SX / DM
# ---------------------------------------------------------------------------------------------------------------------
def _as_matrix(v):
    """Anything -> 2-D object array of S (column vector for 1-D input, CasADi's convention)."""
    if isinstance(v, SX):
        return v.m
    if isinstance(v, S):
        m = np.empty((1, 1), object)
        m[0, 0] = v
        return m
    if isinstance(v, DM):
        v = v.a
    arr = np.asarray(v, dtype=float)
    if arr.ndim == 0:
        arr = arr.reshape(1, 1)
    elif arr.ndim == 1:
        arr = arr.reshape(-1, 1)
    m = np.empty(arr.shape, object)
    for idx in np.ndindex(arr.shape):
        m[idx] = S.of(arr[idx])
    return m


def _bcast(f, x, y):
    a, b = _as_matrix(x), _as_matrix(y)
    if a.shape != b.shape:
        if a.size == 1:
            a = np.broadcast_to(a, b.shape)
        elif b.size == 1:
            b = np.broadcast_to(b, a.shape)
        else:
            raise ValueError(f"shape mismatch {a.shape} vs {b.shape}")
    out = np.empty(a.shape, object)
    for idx in np.ndindex(a.shape):
        out[idx] = f(a[idx], b[idx])
    return SX(out)


class SX:
    """Dense symbolic matrix (2-D object array of scalar nodes), column-major like CasADi."""
    __array_ufunc__ = None

    def __init__(self, m=None, ncol=None):
        if m is None:
            m = np.empty((0, 0), object)
        elif isinstance(m, (int, np.integer)) and ncol is not None:      # SX(n, m): structural zeros
            z = np.empty((int(m), int(ncol)), object)
            for idx in np.ndindex(z.shape):
                z[idx] = S.of(0.0)
            m = z
        elif not (isinstance(m, np.ndarray) and m.dtype == object):
            m = _as_matrix(m)
        self.m = m

    @staticmethod
    def sym(name, n=1, ncol=1):
        m = np.empty((n, ncol), object)
        for j in range(ncol):                      # column-major numbering, as CasADi names X_0, X_1, ...
            for i in range(n):
                m[i, j] = S(G.sym(f"{name}_{j * n + i}"))
        return SX(m)

    @staticmethod
    def zeros(n, ncol=1):
        return SX(int(n), int(ncol))

    @property
    def shape(self):
        return self.m.shape

    @property
    def T(self):
        return SX(self.m.T.copy())

    def size1(self): return self.m.shape[0]
    def size2(self): return self.m.shape[1]
    def numel(self): return self.m.size
    def is_scalar(self): return self.m.size == 1

    def __len__(self):
        return self.m.shape[0]

    def __getitem__(self, k):
        if isinstance(k, tuple):
            r, c = k
            sub = self.m[r if isinstance(r, slice) else [r] if np.isscalar(r) else r][:, c if isinstance(c, slice) else [c] if np.isscalar(c) else c]
            return SX(np.array(sub, dtype=object).reshape(sub.shape))
        flat = self.m.reshape(-1, order="F")
        sub = flat[k]
        if isinstance(sub, S):
            return SX(_as_matrix(sub))
        return SX(np.array(sub, dtype=object).reshape(-1, 1))

    def __iter__(self):
        raise TypeError("SX is not iterable (use vertsplit)")

    def __add__(self, o): return _bcast(lambda a, b: a + b, self, o)
    def __radd__(self, o): return _bcast(lambda a, b: a + b, o, self)
    def __sub__(self, o): return _bcast(lambda a, b: a - b, self, o)
    def __rsub__(self, o): return _bcast(lambda a, b: a - b, o, self)
    def __mul__(self, o): return _bcast(lambda a, b: a * b, self, o)
    def __rmul__(self, o): return _bcast(lambda a, b: a * b, o, self)
    def __truediv__(self, o): return _bcast(lambda a, b: a / b, self, o)
    def __rtruediv__(self, o): return _bcast(lambda a, b: a / b, o, self)
    def __neg__(self): return _map(lambda s: -s, self)
    def __pos__(self): return self
    def __pow__(self, e): return _map(lambda s: s ** e, self)

    def __matmul__(self, o): return mtimes(self, o)
    def __rmatmul__(self, o): return mtimes(o, self)

    def __float__(self):
        c = S.of(self).cv()
        if c is None:
            raise TypeError("symbolic SX has no float value")
        return c

    def __repr__(self):
        return f"SX({self.m.shape[0]}x{self.m.shape[1]})"


MX = SX      # the reference's arm worker uses MX for the same purpose (arm.py:339)


class DM:
    """Numeric matrix returned by Function / solver calls (``.full()`` -> ndarray)."""
    __array_ufunc__ = None

    def __init__(self, a=0.0):
        if isinstance(a, DM):
            a = a.a
        a = np.array(a, dtype=float)
        if a.ndim == 0:
            a = a.reshape(1, 1)
        elif a.ndim == 1:
            a = a.reshape(-1, 1)
        self.a = a

    def full(self): return self.a.copy()
    def toarray(self): return self.a.copy()
    @property
    def shape(self): return self.a.shape
    @property
    def T(self): return DM(self.a.T)
    def __float__(self): return float(self.a.reshape(()))
    def __array__(self, dtype=None, copy=None): return self.a.astype(dtype) if dtype else self.a
    def __getitem__(self, k):
        return DM(self.a[k] if isinstance(k, tuple) else self.a.reshape(-1, order="F")[k])
    def __repr__(self): return f"DM({self.a!r})"
    # arithmetic against SX goes symbolic, against numbers stays numeric
    def _op(self, o, f, swap=False):
        if isinstance(o, (SX, S)):
            return f(o, SX(self.a)) if swap else f(SX(self.a), o)
        ob = o.a if isinstance(o, DM) else np.asarray(o, float)
        return DM(f(ob, self.a) if swap else f(self.a, ob))
    def __add__(self, o): return self._op(o, lambda a, b: a + b)
    def __radd__(self, o): return self._op(o, lambda a, b: a + b, True)
    def __sub__(self, o): return self._op(o, lambda a, b: a - b)
    def __rsub__(self, o): return self._op(o, lambda a, b: a - b, True)
    def __mul__(self, o): return self._op(o, lambda a, b: a * b)
    def __rmul__(self, o): return self._op(o, lambda a, b: a * b, True)
    def __truediv__(self, o): return self._op(o, lambda a, b: a / b)
    def __neg__(self): return DM(-self.a)
    def __matmul__(self, o): return mtimes(self, o)
    def __rmatmul__(self, o): return mtimes(o, self)


def _map(f, x):
    a = _as_matrix(x)
    out = np.empty(a.shape, object)
    for idx in np.ndindex(a.shape):
        out[idx] = f(a[idx])
    return SX(out)


def _numeric(x):
    return not isinstance(x, (SX, S))


def _unary(op, npf):
    def f(x):
        if _numeric(x):
            r = npf(x.a if isinstance(x, DM) else np.asarray(x, float))
            return DM(r) if isinstance(x, DM) else r
        return _map(lambda s: s.un(op), x)
    return f


sin = _unary(_t.SIN, np.sin)
cos = _unary(_t.COS, np.cos)
tanh = _unary(_t.TANH, np.tanh)
exp = _unary(_t.EXP, np.exp)
fabs = _unary(_t.FABS, np.abs)
sqrt = _unary(_t.SQRT, np.sqrt)
log = _unary(_t.LOG, np.log)


def vertcat(*args):
    if not args:
        return SX()
    if all(_numeric(a) for a in args):
        return DM(np.concatenate([_np2(a) for a in args], axis=0))
    mats = [_as_matrix(a) for a in args]
    mats = [m for m in mats if m.size or len(mats) == 1]
    return SX(np.concatenate(mats, axis=0))


def horzcat(*args):
    if all(_numeric(a) for a in args):
        return DM(np.concatenate([_np2(a) for a in args], axis=1))
    return SX(np.concatenate([_as_matrix(a) for a in args], axis=1))


def _np2(a):
    return DM(a).a


def vertsplit(x, incr=1):
    m = _as_matrix(x)
    return [SX(m[i:i + incr].copy()) for i in range(0, m.shape[0], incr)]


def reshape(x, *shape):
    if len(shape) == 1:
        shape = tuple(shape[0])
    if _numeric(x):
        return DM(np.reshape(_np2(x), shape, order="F"))
    return SX(np.reshape(_as_matrix(x), shape, order="F"))


def vec(x):
    return reshape(x, -1, 1)


def sumsqr(x):
    if _numeric(x):
        return DM(np.sum(_np2(x) ** 2))
    r = S.of(0.0)
    for s in _as_matrix(x).reshape(-1, order="F"):
        r = r + s.un(_t.SQ)
    return SX(_as_matrix(r))


def sum1(x):
    m = _as_matrix(x)
    out = np.empty((1, m.shape[1]), object)
    for j in range(m.shape[1]):
        r = S.of(0.0)
        for i in range(m.shape[0]):
            r = r + m[i, j]
        out[0, j] = r
    return SX(out)


def dot(x, y):
    if _numeric(x) and _numeric(y):
        return DM(np.sum(_np2(x) * _np2(y)))
    a, b = _as_matrix(x).reshape(-1, order="F"), _as_matrix(y).reshape(-1, order="F")
    if len(a) != len(b):
        raise ValueError("dot: size mismatch")
    r = S.of(0.0)
    for p, q in zip(a, b):
        r = r + p * q
    return SX(_as_matrix(r))


def mtimes(*args):
    if len(args) == 1 and isinstance(args[0], (list, tuple)):
        args = tuple(args[0])
    r = args[0]
    for nxt in args[1:]:
        r = _mtimes2(r, nxt)
    return r


def _mtimes2(x, y):
    if _numeric(x) and _numeric(y):
        return DM(_np2(x) @ _np2(y))
    a, b = _as_matrix(x), _as_matrix(y)
    if a.size == 1 or b.size == 1:
        return _bcast(lambda p, q: p * q, SX(a), SX(b))
    if a.shape[1] != b.shape[0]:
        raise ValueError(f"mtimes: {a.shape} x {b.shape}")
    out = np.empty((a.shape[0], b.shape[1]), object)
    for i in range(a.shape[0]):
        for j in range(b.shape[1]):
            r = S.of(0.0)
            for k in range(a.shape[1]):
                r = r + a[i, k] * b[k, j]
            out[i, j] = r
    return SX(out)


def diag(x):
    """Vector -> diagonal matrix (structural zeros elsewhere); square matrix -> its diagonal."""
    if _numeric(x):
        a = _np2(x)
        return DM(np.diag(a.reshape(-1)) if 1 in a.shape else np.diag(a).reshape(-1, 1))
    m = _as_matrix(x)
    if 1 in m.shape:
        v = m.reshape(-1, order="F")
        out = np.empty((len(v), len(v)), object)
        for i in range(len(v)):
            for j in range(len(v)):
                out[i, j] = v[i] if i == j else S.of(0.0)
        return SX(out)
    return SX(np.array([m[i, i] for i in range(m.shape[0])], dtype=object).reshape(-1, 1))


def inv(x):
    """Matrix inverse.  Structurally diagonal matrices (the only case on the hot path, rlmpc2.py:409) invert entry-wise;
    anything else by Gauss-Jordan elimination without pivoting."""
    if _numeric(x):
        return DM(np.linalg.inv(_np2(x)))
    m = _as_matrix(x)
    n = m.shape[0]
    if m.shape != (n, n):
        raise ValueError("inv: not square")
    offdiag_zero = all(m[i, j].cv() == 0.0 for i in range(n) for j in range(n) if i != j)
    out = np.empty((n, n), object)
    if offdiag_zero:
        for i in range(n):
            for j in range(n):
                out[i, j] = (1.0 / m[i, i]) if i == j else S.of(0.0)
        return SX(out)
    a = [[m[i, j] for j in range(n)] + [S.of(1.0 if i == j else 0.0) for j in range(n)] for i in range(n)]
    for c in range(n):
        p = a[c][c]
        a[c] = [v / p for v in a[c]]
        for r in range(n):
            if r != c:
                f = a[r][c]
                a[r] = [vr - f * vc for vr, vc in zip(a[r], a[c])]
    for i in range(n):
        for j in range(n):
            out[i, j] = a[i][n + j]
    return SX(out)


def transpose(x):
    return x.T


def norm_2(x):
    return sqrt(sumsqr(x))


def fmin(x, y):
    raise NotImplementedError("fmin is not used on the hot path")


# ---------------------------------------------------------------------------------------------------------------------
# Function / nlpsol
# ---------------------------------------------------------------------------------------------------------------------
def make_tape(inputs, outputs, meta=None):
    """inputs/outputs: {name: SX}.  Extract the sub-graph the outputs depend on as a Tape."""
    in_ids = {k: [s.i for s in _as_matrix(v).reshape(-1, order="F")] for k, v in inputs.items()}
    out_ids = {k: [s.i for s in _as_matrix(v).reshape(-1, order="F")] for k, v in outputs.items()}
    for k, ids in in_ids.items():
        for i in ids:
            if G.op[i] != _t.INPUT:
                raise ValueError(f"input {k} is not purely symbolic")
    need = set()
    stack = [i for ids in out_ids.values() for i in ids]
    while stack:
        i = stack.pop()
        if i in need:
            continue
        need.add(i)
        if G.op[i] > _t.INPUT:
            stack.append(G.a[i])
            if G.b[i] >= 0:
                stack.append(G.b[i])
    declared = {i for ids in in_ids.values() for i in ids}
    for i in need:
        if G.op[i] == _t.INPUT and i not in declared:
            raise ValueError(f"free variable {G.names.get(i)} in Function outputs")
    keep = sorted(need | declared)
    remap = {old: new for new, old in enumerate(keep)}
    op = [G.op[i] for i in keep]
    a = [remap[G.a[i]] if G.a[i] >= 0 else -1 for i in keep]
    b = [remap[G.b[i]] if G.b[i] >= 0 else -1 for i in keep]
    c = [G.cval[i] for i in keep]
    return _t.Tape(op, a, b, c, {k: [remap[i] for i in v] for k, v in in_ids.items()},
                   {k: [remap[i] for i in v] for k, v in out_ids.items()}, meta)


class Function:
    def __init__(self, name, ins, outs, *rest):
        self.name = name
        self.ins = [SX(_as_matrix(v)) for v in ins]
        self.outs = [SX(_as_matrix(v)) for v in outs]
        self._tape = None

    def tape(self):
        if self._tape is None:
            self._tape = make_tape({f"i{k}": v for k, v in enumerate(self.ins)},
                                   {f"o{k}": v for k, v in enumerate(self.outs)}, {"name": self.name})
        return self._tape

    def __call__(self, *args):
        if len(args) != len(self.ins):
            raise TypeError(f"{self.name}: expected {len(self.ins)} arguments")
        if any(not _numeric(a) for a in args):
            res = self._substitute([_as_matrix(a) for a in args])
        else:
            t = self.tape()
            feeds = {}
            for k, (a, proto) in enumerate(zip(args, self.ins)):
                v = _np2(a).reshape(-1, order="F")
                if v.size != proto.m.size:
                    raise ValueError(f"{self.name}: argument {k} has {v.size} entries, expected {proto.m.size}")
                feeds[f"i{k}"] = v
            ev = t.eval(**feeds)
            res = [DM(ev[f"o{k}"].reshape(o.m.shape, order="F")) for k, o in enumerate(self.outs)]
        return res[0] if len(res) == 1 else res

    def _substitute(self, args):
        """Re-trace the output graph with the inputs replaced by the given expressions (what embedding a Function
        call into an SX graph amounts to)."""
        env = {}
        for proto, a in zip(self.ins, args):
            pf, af = proto.m.reshape(-1, order="F"), a.reshape(-1, order="F")
            if len(pf) != len(af):
                raise ValueError(f"{self.name}: argument size {len(af)} != {len(pf)}")
            for p, q in zip(pf, af):
                env[p.i] = q
        out_nodes = [s.i for o in self.outs for s in o.m.reshape(-1, order="F")]
        need, stack = set(), list(out_nodes)
        while stack:
            i = stack.pop()
            if i in need or i in env:
                continue
            need.add(i)
            if G.op[i] > _t.INPUT:
                stack.append(G.a[i])
                if G.b[i] >= 0:
                    stack.append(G.b[i])
        for i in sorted(need):
            o = G.op[i]
            if o == _t.CONST or o == _t.INPUT:
                env[i] = S(i)
            elif o == _t.ADD:
                env[i] = env[G.a[i]] + env[G.b[i]]
            elif o == _t.SUB:
                env[i] = env[G.a[i]] - env[G.b[i]]
            elif o == _t.MUL:
                env[i] = env[G.a[i]] * env[G.b[i]]
            elif o == _t.DIV:
                env[i] = env[G.a[i]] / env[G.b[i]]
            elif o == _t.NEG:
                env[i] = -env[G.a[i]]
            else:
                env[i] = env[G.a[i]].un(o)
        res = []
        for o in self.outs:
            m = np.empty(o.m.shape, object)
            for idx in np.ndindex(o.m.shape):
                m[idx] = env[o.m[idx].i]
            res.append(SX(m))
        return res


CAPTURED = []     # every NlpSolver built since import, in order (how a worker-internal NLP gets out)


class NlpSolver:
    def __init__(self, name, plugin, nlp, opts=None):
        self.name, self.plugin, self.opts = name, plugin, dict(opts or {})
        self.x = SX(_as_matrix(nlp["x"]))
        self.p = SX(_as_matrix(nlp["p"])) if "p" in nlp and nlp["p"] is not None else SX.sym("p_unused", 0)
        self.f = SX(_as_matrix(nlp["f"]))
        self.g = SX(_as_matrix(nlp["g"])) if "g" in nlp and nlp["g"] is not None else SX(np.empty((0, 1), object))
        if self.f.m.size != 1:
            raise ValueError("objective must be scalar")
        self.tape = make_tape({"x": self.x, "p": self.p}, {"f": self.f, "g": self.g},
                              {"name": name, "plugin": plugin})
        self.calls = []
        self.last_stats = {}
        CAPTURED.append(self)

    def __call__(self, x0=0.0, p=None, lbx=-inf, ubx=inf, lbg=-inf, ubg=inf, lam_x0=None, lam_g0=None):
        from . import nlp_solve
        nx, ng, npar = self.x.m.size, self.g.m.size, self.p.m.size

        def vecn(v, n):
            a = _np2(v).reshape(-1, order="F").astype(float)
            if a.size == 1 and n != 1:
                a = np.full(n, a[0])
            if a.size != n:
                raise ValueError(f"{self.name}: expected {n} entries, got {a.size}")
            return a

        args = dict(x0=vecn(x0, nx), p=vecn(p if p is not None else np.zeros(npar), npar),
                    lbx=vecn(lbx, nx), ubx=vecn(ubx, nx), lbg=vecn(lbg, ng), ubg=vecn(ubg, ng))
        res = nlp_solve.solve(self.tape, **args)
        self.calls.append(dict(args, **{k: res[k] for k in ("x", "f", "g", "lam_g", "lam_x", "iters", "kkt", "status")}))
        self.last_stats = {"success": res["status"] == 0, "return_status": "Solve_Succeeded" if res["status"] == 0 else "Failed",
                           "iter_count": res["iters"]}
        return {"x": DM(res["x"]), "f": DM(res["f"]), "g": DM(res["g"]), "lam_g": DM(res["lam_g"]),
                "lam_x": DM(res["lam_x"]), "lam_p": DM(np.zeros(npar))}

    def stats(self):
        return dict(self.last_stats)


def nlpsol(name, plugin, nlp, opts=None):
    return NlpSolver(name, plugin, nlp, opts)
