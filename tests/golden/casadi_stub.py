"""A sympy-backed stand-in for the few `casadi` symbols the reference's NLP builders use
(TEST INFRASTRUCTURE, used by tests/golden/make_reference_vectors.py only).

CasADi is not installable here, but `MPC_optimize.optimize_problem`
(PKG/MPC_CBF_optimize_kin.py:136-255, _kin_pre.py, _dyn.py) only needs symbolic scalars, slicing,
+ - * / **, `vertcat`, `mtimes`, `reshape`, `Function` and `nlpsol`.  With this module registered as
`casadi`, the reference's own, unmodified code builds its objective and constraint expressions as
sympy expressions, which the generator evaluates at seeded points: that pins the restated NLP
(oracle/nlp.py) against the reference itself.  `nlpsol` records the problem and the options; it
does not solve anything (IPOPT stays unpinned).
"""
from __future__ import annotations

import numpy as np
import sympy as sp


class SX:
    """dense matrix of sympy expressions with CasADi's indexing conventions (column-major linear index)."""

    __array_ufunc__ = None  # numpy defers to our reflected operators

    def __init__(self, a):
        a = np.asarray(a, dtype=object)
        if a.ndim == 0:
            a = a.reshape(1, 1)
        elif a.ndim == 1:
            a = a.reshape(-1, 1)
        self.a = a

    # ---- construction
    @staticmethod
    def sym(name, n=1, m=1):
        if n == 1 and m == 1:
            return SX(np.array([[sp.Symbol(name, real=True)]], dtype=object))
        out = np.empty((n, m), dtype=object)
        for j in range(m):
            for i in range(n):
                out[i, j] = sp.Symbol(f"{name}_{i}_{j}", real=True)
        return SX(out)

    # ---- shape
    def size(self):
        return self.a.shape

    @property
    def shape(self):
        return self.a.shape

    @property
    def T(self):
        return SX(self.a.T.copy())

    # ---- indexing
    def __getitem__(self, idx):
        if isinstance(idx, tuple):
            r = self.a[idx]
            if not isinstance(r, np.ndarray):
                return SX(np.array([[r]], dtype=object))
            if r.ndim == 1:
                # X[:, i] is a column, X[i, :] a row
                return SX(r.reshape(-1, 1) if isinstance(idx[1], (int, np.integer)) else r.reshape(1, -1))
            return SX(r)
        flat = self.a.reshape(-1, order="F")[idx]
        return SX(np.atleast_1d(flat).reshape(-1, 1))

    # ---- arithmetic (elementwise, 1x1 and 1-D numpy arrays broadcast like CasADi does)
    @staticmethod
    def _arr(o, like):
        if isinstance(o, SX):
            b = o.a
        else:
            b = np.asarray(o, dtype=object)
            if b.ndim == 0:
                b = b.reshape(1, 1)
            elif b.ndim == 1:
                b = b.reshape(-1, 1) if like.shape[1] == 1 else b.reshape(1, -1)
        return b

    def _bin(self, o, fn, reflected=False):
        b = self._arr(o, self.a)
        x, y = (b, self.a) if reflected else (self.a, b)
        x, y = np.broadcast_arrays(x, y)
        out = np.empty(x.shape, dtype=object)
        for i in np.ndindex(x.shape):
            out[i] = fn(sp.sympify(x[i]), sp.sympify(y[i]))
        return SX(out)

    def __add__(self, o): return self._bin(o, lambda p, q: p + q)
    def __radd__(self, o): return self._bin(o, lambda p, q: p + q, True)
    def __sub__(self, o): return self._bin(o, lambda p, q: p - q)
    def __rsub__(self, o): return self._bin(o, lambda p, q: p - q, True)
    def __mul__(self, o): return self._bin(o, lambda p, q: p * q)
    def __rmul__(self, o): return self._bin(o, lambda p, q: p * q, True)
    def __truediv__(self, o): return self._bin(o, lambda p, q: p / q)
    def __rtruediv__(self, o): return self._bin(o, lambda p, q: p / q, True)
    def __pow__(self, o): return self._bin(o, lambda p, q: p ** q)
    def __neg__(self): return SX(-self.a)

    def map(self, fn):
        out = np.empty(self.a.shape, dtype=object)
        for i in np.ndindex(self.a.shape):
            out[i] = fn(self.a[i])
        return SX(out)


def _sx(o):
    return o if isinstance(o, SX) else SX(np.asarray(o, dtype=object))


def vertcat(*args):
    return SX(np.concatenate([_sx(a).a for a in args], axis=0))


def horzcat(*args):
    return SX(np.concatenate([_sx(a).a for a in args], axis=1))


def reshape(x, n, m):
    return SX(_sx(x).a.reshape((n, m), order="F"))


def mtimes(*args):
    mats = args[0] if len(args) == 1 and isinstance(args[0], (list, tuple)) else args
    out = None
    for m in mats:
        b = m.a if isinstance(m, SX) else np.asarray(m, dtype=object)
        if b.ndim == 1:
            b = b.reshape(-1, 1)
        out = b if out is None else out.dot(b)
    return SX(out)


def cos(x): return _sx(x).map(sp.cos)
def sin(x): return _sx(x).map(sp.sin)
def tan(x): return _sx(x).map(sp.tan)
def sqrt(x): return _sx(x).map(sp.sqrt)
def atan(x): return _sx(x).map(sp.atan)
def fabs(x): return _sx(x).map(sp.Abs)


class DM:
    def __init__(self, a):
        self._a = np.asarray(a, dtype=float)
        if self._a.ndim == 1:
            self._a = self._a.reshape(-1, 1)

    def full(self):
        return self._a.copy()


class Function:
    """ca.Function(name, [inputs], [outputs], in_names, out_names): symbolic or numeric call."""

    def __init__(self, name, ins, outs, in_names=None, out_names=None):
        self.name, self.ins, self.outs = name, [_sx(i) for i in ins], [_sx(o) for o in outs]

    def __call__(self, *args):
        sub = {}
        numeric = True
        for s, a in zip(self.ins, args):
            b = _sx(a).a.reshape(-1, order="F") if isinstance(a, SX) else np.asarray(a, dtype=object).reshape(-1)
            numeric = numeric and not isinstance(a, SX)
            for sym, val in zip(s.a.reshape(-1, order="F"), b):
                sub[sym] = val
        out = self.outs[0].map(lambda e: sp.sympify(e).xreplace(sub))
        if numeric:
            return DM(np.array([[float(v) for v in row] for row in out.a]))
        return out


class _Recorded:
    """what nlpsol was given; calling it is an error (nothing here solves NLPs)"""

    def __init__(self, name, plugin, prob, opts):
        self.name, self.plugin, self.prob, self.opts = name, plugin, prob, dict(opts or {})

    def __call__(self, *a, **k):
        raise RuntimeError("casadi_stub records the NLP; it does not solve it")


def nlpsol(name, plugin, prob, opts=None):
    return _Recorded(name, plugin, prob, opts)
