"""A tiny sympy-backed stand-in for the ~25 ``casadi`` names the reference uses.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Real CasADi 3.7.0 (poetry.lock:90-91 of the
reference) is not installable in this sandbox.  With this module first on ``sys.path`` the
reference's own ``core/runner.py``, ``core/dynamics.py``, ``core/geometry.py``,
``core/utils.py`` and ``core/sdf/*.py`` execute UNMODIFIED and ``Opti`` records the ordered
constraint list and the objective as sympy expressions.  ``oracle/make_golden.py`` lambdifies
those to produce the golden vectors the numpy oracle and the CUDA path are checked against.

Canonical form of ``subject_to`` (CasADi ``OptiNode::canon_expr`` from
memory): if one side has no free symbols it becomes the bound and the other side the row of g;
otherwise g = lhs - rhs with zero bound.
"""
from __future__ import annotations

import math

import numpy as np
import sympy as sp

pi = math.pi
inf = math.inf


def _is_num(x):
    return isinstance(x, (int, float, np.integer, np.floating))


class MX:
    __array_priority__ = 1000
    __hash__ = object.__hash__

    def __init__(self, *args):
        if len(args) == 1:
            a = args[0]
            if isinstance(a, MX):
                self.m = a.m.copy()
            elif isinstance(a, sp.MatrixBase):
                self.m = sp.Matrix(a)
            elif _is_num(a) or isinstance(a, sp.Expr):
                self.m = sp.Matrix([[a]])
            else:  # list / ndarray -> column vector (casadi semantics for 1-D), matrix for 2-D
                arr = np.asarray(a, dtype=object)
                if arr.ndim == 1:
                    arr = arr.reshape(-1, 1)
                self.m = sp.Matrix(arr.shape[0], arr.shape[1], [sp.Float(float(v)) if _is_num(v) else v for v in arr.ravel()])
        elif len(args) == 2:
            self.m = sp.zeros(int(args[0]), int(args[1]))
        else:
            raise TypeError("MX(...)")

    # -- shape ---------------------------------------------------------------------------------
    @property
    def shape(self):
        return (self.m.rows, self.m.cols)

    def size1(self):
        return self.m.rows

    def size2(self):
        return self.m.cols

    def numel(self):
        return self.m.rows * self.m.cols

    @property
    def T(self):
        return MX(self.m.T)

    def is_constant(self):
        return len(self.m.free_symbols) == 0

    def scalar(self):
        assert self.shape == (1, 1), self.shape
        return self.m[0, 0]

    # -- indexing (column-major linear index like casadi) -----------------------------------------
    def _norm(self, idx, n):
        if isinstance(idx, slice):
            return list(range(*idx.indices(n)))
        idx = int(idx)
        return [idx + n if idx < 0 else idx]

    def __getitem__(self, key):
        if isinstance(key, tuple):
            r = self._norm(key[0], self.m.rows)
            c = self._norm(key[1], self.m.cols)
            return MX(self.m.extract(r, c))
        lin = self._norm(key, self.numel())
        flat = [self.m[i % self.m.rows, i // self.m.rows] for i in lin]
        return MX(sp.Matrix(len(flat), 1, flat))

    # -- arithmetic (elementwise, scalars broadcast) ---------------------------------------------
    @staticmethod
    def _coerce(o):
        if isinstance(o, MX):
            return o
        if _is_num(o):
            return MX(float(o))
        if isinstance(o, np.ndarray) and o.ndim == 0:
            return MX(float(o))
        return NotImplemented

    def _bin(self, o, f):
        o = MX._coerce(o)
        if o is NotImplemented:
            return NotImplemented
        a, b = self.m, o.m
        if a.shape == b.shape:
            return MX(sp.Matrix(a.rows, a.cols, [f(x, y) for x, y in zip(a, b)]))
        if a.shape == (1, 1):
            return MX(sp.Matrix(b.rows, b.cols, [f(a[0, 0], y) for y in b]))
        if b.shape == (1, 1):
            return MX(sp.Matrix(a.rows, a.cols, [f(x, b[0, 0]) for x in a]))
        raise ValueError(f"shape mismatch {a.shape} vs {b.shape}")

    def __add__(self, o): return self._bin(o, lambda x, y: x + y)
    def __radd__(self, o): return self._bin(o, lambda x, y: y + x)
    def __sub__(self, o): return self._bin(o, lambda x, y: x - y)
    def __rsub__(self, o): return self._bin(o, lambda x, y: y - x)
    def __mul__(self, o): return self._bin(o, lambda x, y: x * y)
    def __rmul__(self, o): return self._bin(o, lambda x, y: y * x)
    def __truediv__(self, o): return self._bin(o, lambda x, y: x / y)
    def __rtruediv__(self, o): return self._bin(o, lambda x, y: y / x)
    def __pow__(self, o): return self._bin(o, lambda x, y: x ** (int(y) if float(y).is_integer() else y))
    def __neg__(self): return MX(-self.m)

    # -- numpy interop: the reference calls np.sqrt(...) on MX (core/sdf/casadi.py:37) -------------
    def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):
        if method != "__call__":
            return NotImplemented
        name = ufunc.__name__
        una = {"sqrt": sp.sqrt, "exp": sp.exp, "log": sp.log, "sin": sp.sin, "cos": sp.cos, "tan": sp.tan,
               "tanh": sp.tanh, "absolute": sp.Abs, "negative": lambda v: -v}
        if name in una and len(inputs) == 1:
            return MX(self.m.applyfunc(una[name]))
        bina = {"add": lambda x, y: x + y, "subtract": lambda x, y: x - y, "multiply": lambda x, y: x * y,
                "true_divide": lambda x, y: x / y, "power": lambda x, y: x ** y}
        if name in bina and len(inputs) == 2:
            a, b = inputs
            a = a if isinstance(a, MX) else MX._coerce(a)
            if a is NotImplemented:
                return NotImplemented
            return a._bin(b, bina[name])
        return NotImplemented

    # -- relations -> constraint records -----------------------------------------------------------
    def __eq__(self, o): return _Rel("==", self, MX._coerce(o))
    def __ge__(self, o): return _Rel(">=", self, MX._coerce(o))
    def __le__(self, o): return _Rel("<=", self, MX._coerce(o))

    def __repr__(self):
        return f"MX{self.shape}"


SX = MX


class _Rel:
    def __init__(self, op, lhs, rhs, lo=None, hi=None):
        self.op, self.lhs, self.rhs, self.lo, self.hi = op, lhs, rhs, lo, hi


def _ew(f_sym, f_num):
    def fn(x):
        if isinstance(x, MX):
            return MX(x.m.applyfunc(f_sym))
        return f_num(x)
    return fn


cos = _ew(sp.cos, np.cos)
sin = _ew(sp.sin, np.sin)
tan = _ew(sp.tan, np.tan)
sqrt = _ew(sp.sqrt, np.sqrt)
log = _ew(sp.log, np.log)
exp = _ew(sp.exp, np.exp)
tanh = _ew(sp.tanh, np.tanh)
fabs = _ew(sp.Abs, np.abs)


def fmax(a, b):
    if isinstance(a, MX) or isinstance(b, MX):
        return (a if isinstance(a, MX) else MX(a))._bin(b, lambda x, y: sp.Max(x, y))
    return np.maximum(a, b)


def fmin(a, b):
    if isinstance(a, MX) or isinstance(b, MX):
        return (a if isinstance(a, MX) else MX(a))._bin(b, lambda x, y: sp.Min(x, y))
    return np.minimum(a, b)


def vertcat(*args):
    mats = [MX._coerce(a).m for a in args]
    return MX(sp.Matrix.vstack(*mats))


def horzcat(*args):
    mats = [MX._coerce(a).m for a in args]
    return MX(sp.Matrix.hstack(*mats))


def hcat(lst):
    return horzcat(*lst)


def vcat(lst):
    return vertcat(*lst)


def reshape(x, r, c):
    n = x.numel()
    if r == -1:
        r = n // c
    if c == -1:
        c = n // r
    flat = [x.m[i % x.m.rows, i // x.m.rows] for i in range(n)]  # column-major
    out = sp.zeros(r, c)
    for i, v in enumerate(flat):
        out[i % r, i // r] = v
    return MX(out)


def sum1(x):
    return MX(sp.Matrix(1, x.m.cols, [sum(x.m[:, j]) for j in range(x.m.cols)]))


def sum2(x):
    return MX(sp.Matrix(x.m.rows, 1, [sum(x.m[i, :]) for i in range(x.m.rows)]))


def sumsqr(x):
    return MX(sum(v * v for v in x.m))


def mtimes(a, b):
    return MX(MX._coerce(a).m * MX._coerce(b).m)


class _Sol:
    def __init__(self, opti):
        self._o = opti

    def value(self, x):
        return self._o._value(x)


class Opti:
    """Records variables, constraints (canonicalised) and the objective."""

    def __init__(self):
        self.vars = []          # list of MX blocks in creation order
        self.g_rows = []        # sympy expressions, one per scalar row
        self.lbg = []
        self.ubg = []
        self.f = None
        self.initial = {}
        self.solver_name = None
        self.solver_opts = None
        self.debug = self

    def variable(self, n=1, m=1):
        idx = len(self.vars)
        syms = sp.Matrix(n, m, lambda i, j: sp.Symbol(f"v{idx}_{i}_{j}", real=True))
        v = MX(syms)
        self.vars.append(v)
        return v

    def parameter(self, n=1, m=1):
        raise NotImplementedError("the reference does not use Opti.parameter")

    def bounded(self, lo, expr, hi):
        return _Rel("bounded", expr, None, lo, hi)

    def _push(self, e, lo, hi):
        self.g_rows.append(e)
        self.lbg.append(lo)
        self.ubg.append(hi)

    def subject_to(self, rel):
        assert isinstance(rel, _Rel), type(rel)
        if rel.op == "bounded":
            e = rel.lhs
            for i in range(e.numel()):
                self._push(e.m[i % e.m.rows, i // e.m.rows], float(rel.lo), float(rel.hi))
            return
        lhs, rhs = rel.lhs, rel.rhs
        n = max(lhs.numel(), rhs.numel())
        get = lambda x, i: x.m[0, 0] if x.numel() == 1 else x.m[i % x.m.rows, i // x.m.rows]
        for i in range(n):
            a, b = get(lhs, i), get(rhs, i)
            a_const, b_const = len(a.free_symbols) == 0, len(b.free_symbols) == 0
            if rel.op == "==":
                if b_const:
                    self._push(a, float(b), float(b))
                elif a_const:
                    self._push(b, float(a), float(a))
                else:
                    self._push(a - b, 0.0, 0.0)
            elif rel.op == ">=":
                if b_const:
                    self._push(a, float(b), inf)
                elif a_const:
                    self._push(b, -inf, float(a))
                else:
                    self._push(a - b, 0.0, inf)
            elif rel.op == "<=":
                if b_const:
                    self._push(a, -inf, float(b))
                elif a_const:
                    self._push(b, float(a), inf)
                else:
                    self._push(b - a, 0.0, inf)

    def minimize(self, f):
        self.f = MX._coerce(f).scalar()

    def set_initial(self, var, val):
        self.initial[id(var)] = (var, np.asarray(val, dtype=float))

    def solver(self, name, opts=None, *a):
        self.solver_name, self.solver_opts = name, opts

    def solve(self):
        return _Sol(self)

    # decision vector: variables in creation order, each vec'd column-major
    def w_symbols(self):
        out = []
        for v in self.vars:
            out += [v.m[i % v.m.rows, i // v.m.rows] for i in range(v.numel())]
        return out

    def w_initial(self):
        out = []
        for v in self.vars:
            if id(v) in self.initial:
                val = self.initial[id(v)][1].reshape(v.shape)
                out += list(val.ravel(order="F"))
            else:
                out += [0.0] * v.numel()
        return np.array(out, dtype=float)

    def _value(self, x):
        w = dict(zip(self.w_symbols(), self.w_initial()))
        if isinstance(x, MX):
            arr = np.array(x.m.subs(w).evalf(), dtype=float)
            return arr if arr.shape != (1, 1) else float(arr[0, 0])
        if isinstance(x, sp.Expr):
            return float(x.subs(w).evalf())
        return x

    def value(self, x):
        return self._value(x)
