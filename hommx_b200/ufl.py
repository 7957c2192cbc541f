"""Minimal symbolic layer with UFL's operator names.

The reference hands its coefficient ``A(x, y)``, the stratification Jacobian
``Dtheta_transpose(x)`` and the source ``f(x)`` to UFL/FFCx
(/root/reference/src/hommx/hmm.py:131,198,757,1016).  UFL is not a dependency
here: the same callables are traced with the objects of this module, which
build a small expression DAG that ``hommx_b200.codegen`` turns into a CUDA
``__device__`` coefficient program (evaluated on the fly inside the cell
kernel -- no coefficient arrays are materialised) and that can also be
evaluated with numpy for host-side, non-hot-path uses (macro load vector).

Only the operators the reference's tests and examples use are provided
(enumerated from test/integration/*.py and examples/**/*.py): arithmetic,
``sin cos tan acos asin atan sqrt exp ln`` , ``pi``, ``conditional`` with
``< > <= >=``, ``as_vector as_matrix as_tensor Identity indices transpose``.
"""
from __future__ import annotations

import math

import numpy as np

pi = math.pi

_UNARY = ("neg", "sin", "cos", "tan", "acos", "asin", "atan", "sqrt", "exp", "ln", "abs", "not")
_BINARY = ("add", "sub", "mul", "div", "pow", "lt", "gt", "le", "ge", "eq", "ne", "and", "or", "min", "max")
_CONDITION_OPS = ("lt", "gt", "le", "ge", "eq", "ne", "and", "or", "not")


class Expr:
    """Immutable scalar expression node (hash-consed by structural key)."""

    __slots__ = ("op", "args", "value", "key", "_hash")
    __array_ufunc__ = None  # numpy scalars/arrays defer to our reflected operators

    def __init__(self, op, args=(), value=None):
        self.op = op
        self.args = tuple(args)
        self.value = value
        self.key = (op, value, tuple(a.key for a in self.args))
        self._hash = hash(self.key)

    # -- construction helpers ------------------------------------------------
    @staticmethod
    def const(v):
        return Expr("const", (), float(v))

    @staticmethod
    def wrap(v):
        if isinstance(v, Expr):
            return v
        if isinstance(v, (int, float, np.integer, np.floating)):
            return Expr.const(v)
        raise TypeError(f"cannot use {type(v).__name__} in a coefficient expression")

    def is_const(self):
        return self.op == "const"

    def is_condition(self):
        return self.op in _CONDITION_OPS

    # -- arithmetic with light constant folding ------------------------------
    def _bin(self, op, other, reflected=False):
        other = Expr.wrap(other)
        a, b = (other, self) if reflected else (self, other)
        if a.is_const() and b.is_const():
            x, y = a.value, b.value
            if op == "add":
                return Expr.const(x + y)
            if op == "sub":
                return Expr.const(x - y)
            if op == "mul":
                return Expr.const(x * y)
            if op == "div":
                return Expr.const(x / y)
            if op == "pow":
                return Expr.const(x**y)
        if op == "mul":
            if (a.is_const() and a.value == 0.0) or (b.is_const() and b.value == 0.0):
                return Expr.const(0.0)
            if a.is_const() and a.value == 1.0:
                return b
            if b.is_const() and b.value == 1.0:
                return a
        if op == "add":
            if a.is_const() and a.value == 0.0:
                return b
            if b.is_const() and b.value == 0.0:
                return a
        if op == "sub" and b.is_const() and b.value == 0.0:
            return a
        if op == "div" and b.is_const() and b.value == 1.0:
            return a
        return Expr(op, (a, b))

    def __add__(self, o):
        return self._bin("add", o)

    def __radd__(self, o):
        return self._bin("add", o, True)

    def __sub__(self, o):
        return self._bin("sub", o)

    def __rsub__(self, o):
        return self._bin("sub", o, True)

    def __mul__(self, o):
        if isinstance(o, (Tensor, Labeled)):
            return o.__rmul__(self)
        return self._bin("mul", o)

    def __rmul__(self, o):
        return self._bin("mul", o, True)

    def __truediv__(self, o):
        return self._bin("div", o)

    def __rtruediv__(self, o):
        return self._bin("div", o, True)

    def __pow__(self, o):
        return self._bin("pow", o)

    def __rpow__(self, o):
        return self._bin("pow", o, True)

    def __neg__(self):
        if self.is_const():
            return Expr.const(-self.value)
        return Expr("neg", (self,))

    def __pos__(self):
        return self

    def __lt__(self, o):
        return Expr("lt", (self, Expr.wrap(o)))

    def __gt__(self, o):
        return Expr("gt", (self, Expr.wrap(o)))

    def __le__(self, o):
        return Expr("le", (self, Expr.wrap(o)))

    def __ge__(self, o):
        return Expr("ge", (self, Expr.wrap(o)))

    def __hash__(self):
        return self._hash

    def __eq__(self, o):  # structural equality (use ufl.eq for a symbolic condition)
        return isinstance(o, Expr) and self.key == o.key

    def __bool__(self):
        raise TypeError("symbolic conditions have no truth value; use ufl.conditional")

    def __repr__(self):
        if self.op == "const":
            return repr(self.value)
        if self.op == "sym":
            return f"{self.value[0]}[{self.value[1]}]"
        return f"{self.op}({', '.join(map(repr, self.args))})"


def _unary(op, pyfn):
    def f(a):
        a = Expr.wrap(a)
        if a.is_const():
            return Expr.const(pyfn(a.value))
        return Expr(op, (a,))

    f.__name__ = op
    return f


sin = _unary("sin", math.sin)
cos = _unary("cos", math.cos)
tan = _unary("tan", math.tan)
acos = _unary("acos", math.acos)
asin = _unary("asin", math.asin)
atan = _unary("atan", math.atan)
sqrt = _unary("sqrt", math.sqrt)
exp = _unary("exp", math.exp)
ln = _unary("ln", math.log)


def abs(a):  # noqa: A001 - mirrors ufl.algebra.Abs via builtin name as UFL does
    a = Expr.wrap(a)
    return Expr.const(math.fabs(a.value)) if a.is_const() else Expr("abs", (a,))


def lt(a, b):
    return Expr("lt", (Expr.wrap(a), Expr.wrap(b)))


def gt(a, b):
    return Expr("gt", (Expr.wrap(a), Expr.wrap(b)))


def le(a, b):
    return Expr("le", (Expr.wrap(a), Expr.wrap(b)))


def ge(a, b):
    return Expr("ge", (Expr.wrap(a), Expr.wrap(b)))


def eq(a, b):
    return Expr("eq", (Expr.wrap(a), Expr.wrap(b)))


def ne(a, b):
    return Expr("ne", (Expr.wrap(a), Expr.wrap(b)))


def And(a, b):
    return Expr("and", (a, b))


def Or(a, b):
    return Expr("or", (a, b))


def Not(a):
    return Expr("not", (a,))


def max_value(a, b):
    return Expr("max", (Expr.wrap(a), Expr.wrap(b)))


def min_value(a, b):
    return Expr("min", (Expr.wrap(a), Expr.wrap(b)))


def conditional(cond, a, b):
    if not isinstance(cond, Expr) or not cond.is_condition():
        raise TypeError("first argument of conditional must be a condition (a < b, ...)")
    return Expr("cond", (cond, Expr.wrap(a), Expr.wrap(b)))


class Coordinate:
    """``x`` (macro point, a (3,) constant: hmm.py:190-192) or ``y``
    (``ufl.SpatialCoordinate`` of the micro mesh, hmm.py:186)."""

    def __init__(self, name, dim):
        self.name = name
        self.dim = dim

    def __getitem__(self, k):
        k = int(k)
        if not 0 <= k < self.dim:
            raise IndexError(f"{self.name}[{k}] out of range for shape ({self.dim},)")
        return Expr("sym", (), (self.name, k))

    def __len__(self):
        return self.dim

    @property
    def ufl_shape(self):
        return (self.dim,)


# ----------------------------------------------------------------------------
# tensors and index notation
# ----------------------------------------------------------------------------
class Index:
    _count = 0

    def __init__(self):
        Index._count += 1
        self.id = Index._count

    def __repr__(self):
        return f"i{self.id}"


def indices(n):
    return tuple(Index() for _ in range(n))


def _obj(shape, fill=None):
    a = np.empty(shape, dtype=object)
    if fill is not None:
        for idx in np.ndindex(*shape):
            a[idx] = fill
    return a


class Tensor:
    """Dense tensor of scalar expressions (``ufl.as_vector/as_matrix/as_tensor``)."""

    __array_ufunc__ = None

    def __init__(self, data):
        self.data = data  # numpy object array of Expr

    @property
    def ufl_shape(self):
        return self.data.shape

    @property
    def T(self):
        return transpose(self)

    def __getitem__(self, idx):
        if not isinstance(idx, tuple):
            idx = (idx,)
        if any(isinstance(i, Index) for i in idx):
            if not all(isinstance(i, Index) for i in idx) or len(idx) != self.data.ndim:
                raise NotImplementedError("mixed fixed/free indexing is not supported")
            return Labeled(self.data, idx)
        out = self.data[idx]
        return Tensor(out) if isinstance(out, np.ndarray) else out

    def _map(self, fn):
        out = _obj(self.data.shape)
        for idx in np.ndindex(*self.data.shape):
            out[idx] = fn(self.data[idx])
        return Tensor(out)

    def __neg__(self):
        return self._map(lambda e: -e)

    def __add__(self, o):
        if isinstance(o, Tensor):
            if o.data.shape != self.data.shape:
                raise ValueError("shape mismatch in tensor sum")
            out = _obj(self.data.shape)
            for idx in np.ndindex(*self.data.shape):
                out[idx] = self.data[idx] + o.data[idx]
            return Tensor(out)
        return NotImplemented

    def __sub__(self, o):
        return self + (-o)

    def __mul__(self, o):
        if isinstance(o, Tensor):  # UFL: matrix*matrix, matrix*vector are contractions
            a, b = self.data, o.data
            if a.ndim == 2 and b.ndim in (1, 2) and a.shape[1] == b.shape[0]:
                out_shape = (a.shape[0],) + b.shape[1:]
                out = _obj(out_shape)
                for idx in np.ndindex(*out_shape):
                    acc = Expr.const(0.0)
                    for k in range(a.shape[1]):
                        acc = acc + a[idx[0], k] * b[(k,) + idx[1:]]
                    out[idx] = acc
                return Tensor(out)
            raise ValueError("unsupported tensor product shapes")
        return self._map(lambda e: e * Expr.wrap(o))

    def __rmul__(self, o):
        return self._map(lambda e: Expr.wrap(o) * e)

    def __truediv__(self, o):
        return self._map(lambda e: e / Expr.wrap(o))


class Labeled:
    """Tensor components with free indices (``I[i, j] * I[k, l]`` ...)."""

    __array_ufunc__ = None

    def __init__(self, data, labels):
        self.data = data
        self.labels = tuple(labels)

    def _expand(self, labels):
        src = list(self.labels)
        perm = [src.index(l) for l in labels if l in src]
        d = np.transpose(self.data, perm)
        it = iter(d.shape)
        shape = [next(it) if l in src else 1 for l in labels]
        return d.reshape(shape)

    def __mul__(self, o):
        if isinstance(o, Labeled):
            if set(self.labels) & set(o.labels):
                raise NotImplementedError("contraction over repeated indices is not supported")
            labels = list(self.labels) + list(o.labels)
            a, b = self._expand(labels), o._expand(labels)
            shape = np.broadcast_shapes(a.shape, b.shape)
            out = _obj(shape)
            for idx in np.ndindex(*shape):
                ia = tuple(i if s > 1 else 0 for i, s in zip(idx, a.shape))
                ib = tuple(i if s > 1 else 0 for i, s in zip(idx, b.shape))
                out[idx] = a[ia] * b[ib]
            return Labeled(out, labels)
        o = Expr.wrap(o)
        out = _obj(self.data.shape)
        for idx in np.ndindex(*self.data.shape):
            out[idx] = self.data[idx] * o
        return Labeled(out, self.labels)

    def __rmul__(self, o):
        o = Expr.wrap(o)
        out = _obj(self.data.shape)
        for idx in np.ndindex(*self.data.shape):
            out[idx] = o * self.data[idx]
        return Labeled(out, self.labels)

    def __neg__(self):
        return self * -1.0

    def __add__(self, o):
        if not isinstance(o, Labeled) or set(o.labels) != set(self.labels):
            raise ValueError("free indices must agree in a sum")
        b = o._expand(self.labels)
        out = _obj(self.data.shape)
        for idx in np.ndindex(*self.data.shape):
            out[idx] = self.data[idx] + b[idx]
        return Labeled(out, self.labels)

    def __sub__(self, o):
        return self + (-o)


def Identity(d):
    out = _obj((d, d))
    for i in range(d):
        for j in range(d):
            out[i, j] = Expr.const(1.0 if i == j else 0.0)
    return Tensor(out)


def as_vector(comps):
    out = _obj((len(comps),))
    for i, c in enumerate(comps):
        out[i] = Expr.wrap(c)
    return Tensor(out)


def as_matrix(rows):
    out = _obj((len(rows), len(rows[0])))
    for i, r in enumerate(rows):
        if len(r) != len(rows[0]):
            raise ValueError("ragged matrix")
        for j, c in enumerate(r):
            out[i, j] = Expr.wrap(c)
    return Tensor(out)


def as_tensor(expr, indices=None):
    if indices is None:
        if isinstance(expr, Tensor):
            return expr
        arr = np.array(expr, dtype=object)
        out = _obj(arr.shape)
        for idx in np.ndindex(*arr.shape):
            out[idx] = Expr.wrap(arr[idx])
        return Tensor(out)
    if not isinstance(expr, Labeled):
        raise TypeError("as_tensor(expr, indices) needs an index-notation expression")
    src = list(expr.labels)
    return Tensor(np.transpose(expr.data, [src.index(l) for l in indices]).copy())


def transpose(m):
    if not isinstance(m, Tensor) or m.data.ndim != 2:
        raise TypeError("transpose needs a matrix")
    return Tensor(m.data.T.copy())


# ----------------------------------------------------------------------------
# analysis: dependencies, UFL-style degree estimation, numpy evaluation
# ----------------------------------------------------------------------------
def depends_on(e, name, _memo=None):
    """True if the expression reads coordinate ``name`` ('x' or 'y')."""
    _memo = {} if _memo is None else _memo
    r = _memo.get(e.key)
    if r is None:
        if e.op == "sym":
            r = e.value[0] == name
        else:
            r = any(depends_on(a, name, _memo) for a in e.args)
        _memo[e.key] = r
    return r


def estimate_degree(e, sym_degree):
    """UFL's ``estimate_total_polynomial_degree`` rules for the operators above
    (SURVEY.md A.5): sum -> max, product/division -> sum, math functions ->
    argument + 2, non-negative integer power -> degree * p, conditional ->
    max(true, false) with the condition ignored.  ``sym_degree`` maps 'x'/'y'
    to the degree of that coordinate (0 for the macro constant, 1 for the
    micro SpatialCoordinate on an affine mesh)."""
    op = e.op
    if op == "const":
        return 0
    if op == "sym":
        return sym_degree[e.value[0]]
    d = [estimate_degree(a, sym_degree) for a in e.args]
    if op in ("add", "sub", "min", "max"):
        return max(d)
    if op in ("mul", "div"):
        return d[0] + d[1]
    if op == "neg" or op == "abs":
        return d[0]
    if op == "pow":
        p = e.args[1]
        if p.is_const() and p.value >= 0 and float(p.value).is_integer():
            return d[0] * int(p.value)
        return d[0] + 2
    if op == "cond":
        return max(d[1], d[2])
    if op in _CONDITION_OPS:
        return 0
    return d[0] + 2  # sin, cos, sqrt, exp, ...


_NP = {
    "neg": np.negative, "sin": np.sin, "cos": np.cos, "tan": np.tan, "acos": np.arccos,
    "asin": np.arcsin, "atan": np.arctan, "sqrt": np.sqrt, "exp": np.exp, "ln": np.log,
    "abs": np.abs, "not": np.logical_not, "add": np.add, "sub": np.subtract, "mul": np.multiply,
    "div": np.divide, "pow": np.power, "lt": np.less, "gt": np.greater, "le": np.less_equal,
    "ge": np.greater_equal, "eq": np.equal, "ne": np.not_equal, "and": np.logical_and,
    "or": np.logical_or, "min": np.minimum, "max": np.maximum,
}  # fmt: skip


def evaluate(e, env, _memo=None):
    """Evaluate with numpy; ``env`` maps ('x', k) / ('y', k) to floats or arrays.
    Host-side helper (macro load vector, tests); never used on the hot path."""
    _memo = {} if _memo is None else _memo
    if e.key in _memo:
        return _memo[e.key]
    if e.op == "const":
        r = e.value
    elif e.op == "sym":
        r = env[e.value]
    elif e.op == "cond":
        c, a, b = (evaluate(t, env, _memo) for t in e.args)
        r = np.where(c, a, b)
    else:
        r = _NP[e.op](*(evaluate(t, env, _memo) for t in e.args))
    _memo[e.key] = r
    return r


# ----------------------------------------------------------------------------
# translator from REAL UFL expression DAGs (BASELINE north star: the classes stay drop-ins for scripts that
# `import ufl`; hmm.py:190-198 calls A(fem.Constant, ufl.SpatialCoordinate))
# ----------------------------------------------------------------------------
_UFL_UNARY = {"Sin": "sin", "Cos": "cos", "Tan": "tan", "Acos": "acos", "Asin": "asin", "Atan": "atan", "Sqrt": "sqrt",
              "Exp": "exp", "Ln": "ln", "Abs": "abs"}  # fmt: skip
_UFL_COND = {"LT": "lt", "GT": "gt", "LE": "le", "GE": "ge", "EQ": "eq", "NE": "ne"}


def _index_id(i):
    c = getattr(i, "count", None)
    return c() if callable(c) else (c if c is not None else getattr(i, "_count", id(i)))


def _multiindex(mi):
    it = mi.indices() if callable(getattr(mi, "indices", None)) else getattr(mi, "_indices", mi)
    return list(it)


def from_ufl(expr, terminals):
    """Translate a UFL expression into this module's IR (``Expr`` for a scalar, ``Tensor`` otherwise).

    ``terminals`` lists ``(object, name)``: the objects the callable received -- the macro point (``fem.Constant`` of
    shape (3,), hmm.py:190-192) as ``"x"``, the ``ufl.SpatialCoordinate`` of the micro mesh (hmm.py:186) as ``"y"``;
    every other terminal must be a number.  The walk only uses what every UFL node offers -- the class name, ``ufl_operands``, ``ufl_shape`` -- so
    it needs no ``import ufl`` here (and is exercised in this repository against tests/fake_ufl, a stand-in with
    UFL's class names: UFL itself is not installable in the build image).  Covered: the operators the reference's
    tests and examples use (arithmetic, powers, ``sin cos tan acos asin atan sqrt exp ln abs``, ``conditional`` with
    comparisons and and/or/not, ``min_value max_value``), tensors (``as_vector as_matrix as_tensor Identity
    transpose``), index notation (``Indexed ComponentTensor IndexSum``) and ``dot inner outer``."""

    def shape(n):
        return tuple(getattr(n, "ufl_shape", ()))

    def comp(n, idx, env):
        name = type(n).__name__
        for t, tname in terminals:
            if n is t:
                return Expr("sym", (), (tname, int(idx[0])))
        if isinstance(n, (int, float, np.integer, np.floating)):
            return Expr.const(n)
        ops = getattr(n, "ufl_operands", ())
        if name in ("IntValue", "FloatValue", "RealValue", "ScalarValue", "ComplexValue"):
            v = n.value() if callable(getattr(n, "value", None)) else getattr(n, "_value", None)
            return Expr.const(float(v))
        if name == "Zero":
            return Expr.const(0.0)
        if name == "Identity":
            return Expr.const(1.0 if idx[0] == idx[1] else 0.0)
        if name == "Indexed":
            sub = []
            for i in _multiindex(ops[1]):
                sub.append(int(i) if type(i).__name__ == "FixedIndex" else env[_index_id(i)])
            return comp(ops[0], tuple(sub) + tuple(idx), env)
        if name == "ComponentTensor":
            free = _multiindex(ops[1])
            e2 = dict(env)
            for i, v in zip(free, idx[: len(free)]):
                e2[_index_id(i)] = v
            return comp(ops[0], tuple(idx[len(free):]), e2)
        if name == "IndexSum":
            (i,) = _multiindex(ops[1])
            dim = n.dimension() if callable(getattr(n, "dimension", None)) else n._dimension
            acc = Expr.const(0.0)
            for v in range(int(dim)):
                e2 = dict(env)
                e2[_index_id(i)] = v
                acc = acc + comp(ops[0], idx, e2)
            return acc
        if name == "ListTensor":
            return comp(ops[idx[0]], tuple(idx[1:]), env)
        if name == "Transposed":
            return comp(ops[0], (idx[1], idx[0]), env)
        if name == "Sum":
            return comp(ops[0], idx, env) + comp(ops[1], idx, env)
        if name == "Product":
            out = Expr.const(1.0)
            for o in ops:  # at most one factor is tensor-valued
                out = out * comp(o, idx if shape(o) else (), env)
            return out
        if name == "Division":
            return comp(ops[0], idx, env) / comp(ops[1], (), env)
        if name == "Power":
            return comp(ops[0], (), env) ** comp(ops[1], (), env)
        if name in _UFL_UNARY:
            a = comp(ops[0], (), env)
            return globals()[_UFL_UNARY[name]](a)
        if name in _UFL_COND:
            return Expr(_UFL_COND[name], (comp(ops[0], (), env), comp(ops[1], (), env)))
        if name == "AndCondition":
            return Expr("and", (comp(ops[0], (), env), comp(ops[1], (), env)))
        if name == "OrCondition":
            return Expr("or", (comp(ops[0], (), env), comp(ops[1], (), env)))
        if name == "NotCondition":
            return Expr("not", (comp(ops[0], (), env),))
        if name == "Conditional":
            return conditional(comp(ops[0], (), env), comp(ops[1], idx, env), comp(ops[2], idx, env))
        if name in ("MinValue", "MaxValue"):
            return Expr("min" if name == "MinValue" else "max", (comp(ops[0], (), env), comp(ops[1], (), env)))
        if name in ("Dot", "Inner", "Outer"):
            sa, sb = shape(ops[0]), shape(ops[1])
            if name == "Outer":
                return comp(ops[0], idx[: len(sa)], env) * comp(ops[1], idx[len(sa):], env)
            if name == "Inner":
                acc = Expr.const(0.0)
                for k in np.ndindex(*sa):
                    acc = acc + comp(ops[0], k, env) * comp(ops[1], k, env)
                return acc
            acc = Expr.const(0.0)
            for k in range(sa[-1]):
                acc = acc + comp(ops[0], tuple(idx[: len(sa) - 1]) + (k,), env) * comp(ops[1], (k,) + tuple(idx[len(sa) - 1:]), env)
            return acc
        raise NotImplementedError(f"UFL node {name} is not supported in coefficient expressions")

    if isinstance(expr, (Expr, Tensor)):
        return expr
    if isinstance(expr, (int, float, np.integer, np.floating)):
        return Expr.const(expr)
    if isinstance(expr, (list, tuple)):
        parts = [from_ufl(e, terminals) for e in expr]
        return as_tensor([p.data.tolist() if isinstance(p, Tensor) else p for p in parts])
    sh = shape(expr)
    if not sh:
        return comp(expr, (), {})
    out = _obj(sh)
    for idx in np.ndindex(*sh):
        out[idx] = comp(expr, idx, {})
    return Tensor(out)


def is_foreign(val):
    """True for an object that is not part of this module's IR but looks like a UFL expression."""
    if isinstance(val, (Expr, Tensor, int, float, np.integer, np.floating)):
        return False
    return hasattr(val, "ufl_operands") or hasattr(val, "ufl_shape")


def foreign_module(fn):
    """The UFL-like module a user callable is written against (found in the callable's globals and closure: any module
    other than this one that offers ``SpatialCoordinate``), or None."""
    import types

    inner = getattr(fn, "__wrapped__", None)  # PoissonPeriodicHMM wraps A(y) into A(x, y)
    if inner is not None:
        return foreign_module(inner)
    seen = list(getattr(fn, "__globals__", {}).values())
    for cell in getattr(fn, "__closure__", None) or ():
        try:
            seen.append(cell.cell_contents)
        except ValueError:
            pass
    for v in seen:
        if isinstance(v, types.ModuleType) and v.__name__ != __name__ and hasattr(v, "SpatialCoordinate"):
            return v
    return None


def foreign_symbols(mod, dim):
    """``(x_const, coord)`` of the foreign module, built the way the reference builds them: a (3,) ``Constant`` for the
    macro point (hmm.py:190-192) and the ``SpatialCoordinate`` of a ``dim``-dimensional affine simplex mesh
    (hmm.py:130, 186)."""
    cell = {2: "triangle", 3: "tetrahedron"}[dim]
    try:
        import basix.ufl

        element = basix.ufl.element("Lagrange", cell, 1, shape=(dim,))
    except ImportError:
        element = mod.VectorElement("Lagrange", cell, 1) if hasattr(mod, "VectorElement") else cell
    domain = mod.Mesh(element)
    return mod.Constant(domain, shape=(3,)), mod.SpatialCoordinate(domain)


def call_traced(fn, args):
    """Call a user callable (``A(x, y)``, ``Dtheta_transpose(x)``, ``f(x)``) and return its value in this module's IR.

    ``args`` lists what each positional argument is: ``("const", name)`` for the macro point handed over as a (3,)
    constant, ``("coord", name, dim)`` for a spatial coordinate.  The callable is first traced with this module's own
    symbols (scripts written against ``hommx_b200.ufl``); if that raises, or the callable's module globals hold a UFL
    module, it is called with that module's ``Constant`` / ``SpatialCoordinate`` and the resulting UFL DAG is translated
    (``from_ufl``)."""
    own = [Coordinate(a[1], 3 if a[0] == "const" else a[2]) for a in args]
    mod = foreign_module(fn)
    first = None
    if mod is None:
        val = fn(*own)
        if not is_foreign(val):
            return val
        raise TypeError(f"the callable returned a {type(val).__name__}: write it against hommx_b200.ufl or UFL")
    try:
        val = fn(*own)
        if not is_foreign(val) and not (isinstance(val, (list, tuple)) and any(is_foreign(v) for v in val)):
            return val
    except Exception as e:  # UFL rejects this module's symbols (as_ufl): trace with its own
        first = e
    syms, terminals = [], []
    for a in args:
        dim = 3 if a[0] == "const" else a[2]
        xc, yc = foreign_symbols(mod, dim)
        s = xc if a[0] == "const" else yc
        syms.append(s)
        terminals.append((s, a[1]))
    try:
        val = fn(*syms)
    except Exception as e:
        raise e from first
    return from_ufl(val, terminals)
