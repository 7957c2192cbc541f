"""Coefficient front-end: trace ``A(x, y)`` / ``Dtheta_transpose(x)`` and emit
the CUDA coefficient program the cell kernel evaluates on the fly.

Replaces, for the hot path, what the reference gets from UFL + FFCx JIT:
``self._A_micro = self._coeff(self._x_macro, self._y)`` and the
``1 + n_b + n_b^2`` ``fem.form`` compilations
(/root/reference/src/hommx/hmm.py:198, 259-274, 306, 756-757, 1015-1016).

Because P1 gradients are constant per micro element, the operator depends on
the coefficient only through its per-element quadrature mean.
The tracer therefore splits every tensor component into

    A_c(x, y) = c0_c(x) + sum_k c_ck(x) * s_k(x, y)

where the *atoms* ``s_k`` are the distinct y-dependent scalar sub-expressions
(e.g. only ``mu(y)`` for an isotropic Hooke tensor with constant lambda).  The
kernel stores ``NATOMS`` doubles per micro element instead of a full tensor,
and x-only sub-expressions are hoisted into per-point constants ``pc``.
"""
from __future__ import annotations

import hashlib
from dataclasses import dataclass, field

import numpy as np

from . import ufl
from .ufl import Expr, Tensor

POISSON, ELASTICITY = 0, 1


def voigt_pairs(d):
    """Unit-strain basis order: diagonal entries, then (0,1)[,(0,2),(1,2)]."""
    return [(q, q) for q in range(d)] + ([(0, 1)] if d == 2 else [(0, 1), (0, 2), (1, 2)])


@dataclass
class CoefficientProgram:
    dim: int
    kind: int
    stratified: bool
    degree: int  # UFL-estimated quadrature degree of the cell-problem forms
    ydep: int  # bit k set: some atom reads y[k]
    natoms: int
    npc: int
    ncomp: int
    scalar: bool  # Poisson: A is a scalar times the identity
    source: str  # CUDA struct body (deterministic text; hashed by the library)
    atoms: list = field(default_factory=list)
    comps: list = field(default_factory=list)  # traced unique tensor components (Expr)
    dtheta: list | None = None  # traced d*d Expr, row-major M[p][i]

    @property
    def key(self):
        return hashlib.sha1(self.source.encode()).hexdigest()[:16]

    @property
    def n_rhs(self):
        d = self.dim
        return d if self.kind == POISSON else d * (d + 1) // 2

    # host-side evaluation (tests, drop-in classes) ---------------------------
    def eval_dtheta(self, x):
        d = self.dim
        if self.dtheta is None:
            return np.eye(d)
        env = {("x", k): float(x[k]) for k in range(3)}
        return np.array([float(ufl.evaluate(e, env)) for e in self.dtheta]).reshape(d, d)


# ----------------------------------------------------------------------------
# tracing
# ----------------------------------------------------------------------------
def _numeric_equal(a: Expr, b: Expr, dim, rng):
    if a.key == b.key:
        return True
    for _ in range(4):
        env = {("x", k): rng.uniform(0.1, 0.9) for k in range(3)}
        env.update({("y", k): rng.uniform(0.0, 1.0) for k in range(dim)})
        va, vb = float(ufl.evaluate(a, env)), float(ufl.evaluate(b, env))
        if abs(va - vb) > 1e-13 * max(1.0, abs(va), abs(vb)):
            return False
    return True


def _trace_components(A, dim, kind):
    val = ufl.call_traced(A, [("const", "x"), ("coord", "y", dim)])  # hmm.py:198: A(x_macro, y)
    rng = np.random.default_rng(0)
    if kind == POISSON:
        if isinstance(val, Tensor):
            if val.data.shape != (dim, dim):
                raise ValueError(f"A must be a scalar or a {dim}x{dim} matrix, got shape {val.data.shape}")
            for i in range(dim):
                for j in range(i + 1, dim):
                    if not _numeric_equal(val.data[i, j], val.data[j, i], dim, rng):
                        raise NotImplementedError(
                            "non-symmetric diffusion tensors need a non-symmetric Krylov method; "
                            "the CUDA path solves the cell problems with PCG"
                        )
            comps = [val.data[i, j] for i in range(dim) for j in range(i, dim)]
            return comps, False
        a = Expr.wrap(val)
        zero = Expr.const(0.0)
        comps = [a if i == j else zero for i in range(dim) for j in range(i, dim)]
        return comps, True
    if not isinstance(val, Tensor) or val.data.shape != (dim,) * 4:
        raise ValueError(f"elasticity needs a rank-4 tensor of shape {(dim,) * 4}")
    T = val.data
    pairs = voigt_pairs(dim)
    for (i, j) in pairs:
        for (k, l) in pairs:
            ok = (
                _numeric_equal(T[i, j, k, l], T[j, i, k, l], dim, rng)
                and _numeric_equal(T[i, j, k, l], T[i, j, l, k], dim, rng)
                and _numeric_equal(T[i, j, k, l], T[k, l, i, j], dim, rng)
            )
            if not ok:
                raise NotImplementedError("the elasticity tensor must have minor and major symmetries")
    m = len(pairs)
    comps = [T[pairs[p] + pairs[q]] for p in range(m) for q in range(p, m)]
    return comps, False


def _trace_dtheta(Dtheta_transpose, dim):
    M = ufl.call_traced(Dtheta_transpose, [("const", "x")])  # hmm.py:757: Dtheta_transpose(x_macro)
    if not isinstance(M, Tensor) or M.data.shape != (dim, dim):
        shape = getattr(getattr(M, "data", None), "shape", None)
        raise ValueError(
            f"Dtheta_transpose must return a {dim}x{dim} matrix (got {shape}); micro and macro "
            "meshes have equal dimension (hmm.py:114-115)"
        )
    flat = [M.data[p, i] for p in range(dim) for i in range(dim)]
    for e in flat:
        if ufl.depends_on(e, "y"):
            raise ValueError("Dtheta_transpose may only depend on the macro point x")
    return flat


# ----------------------------------------------------------------------------
# affine decomposition in y-dependent atoms
# ----------------------------------------------------------------------------
def _affine(e: Expr, dep):
    """-> (c0: Expr, {atom_key: (atom_expr, coef_expr)}) with x-only c0/coef."""
    if not dep(e):
        return e, {}
    op = e.op
    if op in ("add", "sub"):
        c0a, la = _affine(e.args[0], dep)
        c0b, lb = _affine(e.args[1], dep)
        sgn = 1.0 if op == "add" else -1.0
        out = dict(la)
        for k, (atom, coef) in lb.items():
            coef = coef if sgn > 0 else -coef
            out[k] = (atom, out[k][1] + coef) if k in out else (atom, coef)
        return (c0a + c0b if sgn > 0 else c0a - c0b), out
    if op == "neg":
        c0, l = _affine(e.args[0], dep)
        return -c0, {k: (a, -c) for k, (a, c) in l.items()}
    if op == "mul":
        a, b = e.args
        if not dep(a) or not dep(b):
            s, t = (a, b) if not dep(a) else (b, a)
            c0, l = _affine(t, dep)
            return s * c0, {k: (at, s * c) for k, (at, c) in l.items()}
    if op == "div" and not dep(e.args[1]):
        c0, l = _affine(e.args[0], dep)
        den = e.args[1]
        return c0 / den, {k: (at, c / den) for k, (at, c) in l.items()}
    return Expr.const(0.0), {e.key: (e, Expr.const(1.0))}


# ----------------------------------------------------------------------------
# CUDA emission
# ----------------------------------------------------------------------------
_FUN = {"sin": "sin", "cos": "cos", "tan": "tan", "acos": "acos", "asin": "asin", "atan": "atan",
        "sqrt": "sqrt", "exp": "exp", "ln": "log", "abs": "fabs"}  # fmt: skip
_INFIX = {"add": "+", "sub": "-", "mul": "*", "div": "/", "lt": "<", "gt": ">", "le": "<=", "ge": ">=",
          "eq": "==", "ne": "!=", "and": "&&", "or": "||"}  # fmt: skip


class _Emitter:
    def __init__(self, symmap, prefix):
        self.symmap = symmap  # (name, k) -> C expression
        self.lines = []
        self.memo = {}
        self.prefix = prefix

    def ref(self, e: Expr):
        if e.key in self.memo:
            return self.memo[e.key]
        op = e.op
        if op == "const":
            v = e.value
            r = repr(v) if v >= 0 else f"({v!r})"
            if "e" not in r and "." not in r and "inf" not in r and "nan" not in r:
                r += ".0"
            self.memo[e.key] = r
            return r
        if op == "sym":
            r = self.symmap[e.value]
            self.memo[e.key] = r
            return r
        a = [self.ref(t) for t in e.args]
        if op in _INFIX:
            rhs = f"{a[0]} {_INFIX[op]} {a[1]}"
        elif op in _FUN:
            rhs = f"{_FUN[op]}({a[0]})"
        elif op == "neg":
            rhs = f"-{a[0]}"
        elif op == "not":
            rhs = f"!{a[0]}"
        elif op == "min":
            rhs = f"fmin({a[0]}, {a[1]})"
        elif op == "max":
            rhs = f"fmax({a[0]}, {a[1]})"
        elif op == "cond":
            rhs = f"{a[0]} ? {a[1]} : {a[2]}"
        elif op == "pow":
            p = e.args[1]
            if p.is_const() and float(p.value).is_integer() and 1 <= p.value <= 4:
                rhs = " * ".join([a[0]] * int(p.value))
            else:
                rhs = f"pow({a[0]}, {a[1]})"
        else:
            raise NotImplementedError(op)
        name = f"{self.prefix}{len(self.lines)}"
        ctype = "bool" if e.is_condition() else "double"
        self.lines.append(f"    const {ctype} {name} = {rhs};")
        self.memo[e.key] = name
        return name


def _hoist_point_constants(exprs, pcs):
    """Replace maximal x-only non-constant sub-expressions by pc symbols."""
    memo_dep = {}, {}, {}

    def dep_y(e):  # varies within a point: micro coordinate, averaged atoms, strain
        return any(ufl.depends_on(e, n, m) for n, m in zip(("y", "s", "e"), memo_dep))

    memo_x = {}
    dep_x = lambda e: ufl.depends_on(e, "x", memo_x)  # noqa: E731
    cache = {}

    def visit(e):
        if e.key in cache:
            return cache[e.key]
        if e.op == "const":
            r = e
        elif not dep_y(e) and dep_x(e):
            if e.key not in pcs:
                pcs[e.key] = (len(pcs), e)
            r = Expr("sym", (), ("pc", pcs[e.key][0]))
        elif e.op == "sym":
            r = e
        else:
            r = Expr(e.op, tuple(visit(a) for a in e.args), e.value)
        cache[e.key] = r
        return r

    return [visit(e) for e in exprs]


def _reads(e, sym, _memo=None):
    """True if the expression reads the symbol ``sym`` = (name, k)."""
    _memo = {} if _memo is None else _memo
    r = _memo.get(e.key)
    if r is None:
        r = e.value == sym if e.op == "sym" else any(_reads(a, sym, _memo) for a in e.args)
        _memo[e.key] = r
    return r


def _body(lines, outs, target):
    out = list(lines)
    for i, r in enumerate(outs):
        out.append(f"    {target}[{i}] = {r};")
    return "\n".join(out)


def build_program(A, dim, kind, Dtheta_transpose=None) -> CoefficientProgram:
    """Trace the reference-style callables and emit the coefficient program."""
    if dim not in (2, 3):
        raise ValueError("Topology should be 3D or 2D")  # hmm.py:104-105
    comps, scalar = _trace_components(A, dim, kind)
    degree = max(ufl.estimate_degree(c, {"x": 0, "y": 1}) for c in comps)
    dth = _trace_dtheta(Dtheta_transpose, dim) if Dtheta_transpose is not None else None

    memo_dep = {}
    dep = lambda e: ufl.depends_on(e, "y", memo_dep)  # noqa: E731
    atoms, atom_index = [], {}
    affine = []
    for c in comps:
        c0, lin = _affine(c, dep)
        terms = []
        for k, (atom, coef) in lin.items():
            if k not in atom_index:
                atom_index[k] = len(atoms)
                atoms.append(atom)
            terms.append((atom_index[k], coef))
        affine.append((c0, terms))
    natoms = len(atoms)

    # tensor components as expressions of s[k] with x-only coefficients
    def comp_expr(c0, terms):
        e = c0
        for k, coef in terms:
            e = e + coef * Expr("sym", (), ("s", k))
        return e

    tens = [comp_expr(c0, t) for c0, t in affine]
    pcs = {}
    atoms_h = _hoist_point_constants(atoms, pcs)
    tens_h = _hoist_point_constants(tens, pcs)
    m = len(voigt_pairs(dim))
    sig_exprs = []
    if kind == ELASTICITY:
        idx = {}
        n = 0
        for p in range(m):
            for q in range(p, m):
                idx[(p, q)] = idx[(q, p)] = n
                n += 1
        esym = [Expr("sym", (), ("e", q)) for q in range(m)]
        # If the normal-normal couplings share one value c (isotropic and cubic materials: lambda / C12),
        #   sigma_p = c tr(e) + (T_pp - c) e_p + (normal-shear couplings)      for the normal components,
        # which saves 5 of 14 FP64 operations per simplex for Hooke.  The differences T_pp - c are formed on the
        # affine representation (x-only coefficients fold or become point constants), not per element.
        offdiag = [(p, q) for p in range(dim) for q in range(p + 1, dim)]
        shared = all(affine[idx[pq]][0].key == affine[idx[offdiag[0]]][0].key
                     and {k: c.key for k, c in affine[idx[pq]][1]} == {k: c.key for k, c in affine[idx[offdiag[0]]][1]}
                     for pq in offdiag)  # fmt: skip
        sig_exprs = []
        if shared:
            c0c, tc = affine[idx[offdiag[0]]]
            cdict = dict(tc)
            tr = esym[0]
            for q in range(1, dim):
                tr = tr + esym[q]
            common = tens_h[idx[offdiag[0]]] * tr
            for p in range(m):
                if p < dim:
                    c0p, tp = affine[idx[(p, p)]]
                    diff = c0p - c0c
                    keys = sorted(set(k for k, _ in tp) | set(cdict))
                    tpd = dict(tp)
                    for k in keys:
                        ck = tpd.get(k, Expr.const(0.0)) - cdict.get(k, Expr.const(0.0))
                        diff = diff + ck * Expr("sym", (), ("s", k))
                    diff = _hoist_point_constants([diff], pcs)[0]
                    acc = common + diff * esym[p]
                    for q in range(dim, m):
                        acc = acc + tens_h[idx[(p, q)]] * esym[q]
                else:
                    acc = Expr.const(0.0)
                    for q in range(m):
                        acc = acc + tens_h[idx[(p, q)]] * esym[q]
                sig_exprs.append(acc)
        else:
            for p in range(m):
                acc = Expr.const(0.0)
                for q in range(m):
                    acc = acc + tens_h[idx[(p, q)]] * esym[q]
                sig_exprs.append(acc)
    # affine coefficients: C[0*NCOMP + c] = c0_c(x), C[(1+k)*NCOMP + c] = c_ck(x)
    aff = [c0 for c0, _ in affine]
    for k in range(natoms):
        for c0, terms in affine:
            coef = Expr.const(0.0)
            for kk, cf in terms:
                if kk == k:
                    coef = coef + cf
            aff.append(coef)
    aff_h = _hoist_point_constants(aff, pcs)
    npc = len(pcs)

    sym_x = {("x", k): f"x[{k}]" for k in range(3)}
    sym_y = {("y", k): f"y[{k}]" for k in range(dim)}
    sym_pc = {("pc", k): f"pc[{k}]" for k in range(max(npc, 1))}
    sym_s = {("s", k): f"s[{k}]" for k in range(max(natoms, 1))}

    em = _Emitter(sym_x, "p")
    pc_refs = [em.ref(e) for _, e in sorted(pcs.values(), key=lambda t: t[0])]
    pc_body = _body(em.lines, pc_refs, "pc")

    em = _Emitter({**sym_pc, **sym_y}, "a")
    at_refs = [em.ref(e) for e in atoms_h]
    at_body = _body(em.lines, at_refs, "s")

    em = _Emitter({**sym_pc, **sym_s}, "t")
    te_refs = [em.ref(e) for e in tens_h]
    te_body = _body(em.lines, te_refs, "A")

    em = _Emitter(sym_pc, "c")
    af_body = _body(em.lines, [em.ref(e) for e in aff_h], "C")

    ydep = 0
    for a in atoms:
        for k in range(dim):
            if _reads(a, ("y", k)):
                ydep |= 1 << k

    m = len(voigt_pairs(dim))
    if kind == ELASTICITY:
        sym_e = {("e", k): f"e[{k}]" for k in range(m)}
        em = _Emitter({**sym_pc, **sym_s, **sym_e}, "g")
        sig = [em.ref(e) for e in sig_exprs]
        st_body = _body(em.lines, sig, "sig")
    else:
        st_body = "    (void)pc; (void)s; (void)e; (void)sig;"

    if dth is not None:
        em = _Emitter(sym_x, "m")
        dt_body = _body(em.lines, [em.ref(e) for e in dth], "M")
    else:
        dt_body = "\n".join(
            f"    M[{p * dim + i}] = {1.0 if p == i else 0.0};" for p in range(dim) for i in range(dim)
        )

    # structurally zero entries of M (bit p*dim+i): the kernels drop those terms at compile time
    if dth is not None:
        mzero = sum(1 << k for k, e in enumerate(dth) if e.is_const() and e.value == 0.0)
    else:
        mzero = sum(1 << (p * dim + i) for p in range(dim) for i in range(dim) if p != i)

    fn = "  __device__ __forceinline__ static void"
    src = f"""// hommx_b200 coefficient program v1
struct HMX_COEFF {{
  static constexpr int DIM = {dim};
  static constexpr int KIND = {kind};
  static constexpr int STRATIFIED = {1 if dth is not None else 0};
  static constexpr int NATOMS = {natoms};
  static constexpr int NPC = {npc};
  static constexpr int NCOMP = {len(comps)};
  static constexpr int SCALAR = {1 if scalar else 0};
  static constexpr int QDEG = {degree};
  static constexpr int YDEP = {ydep};
  static constexpr unsigned MZERO = {mzero}u;
{fn} point_consts(const double* __restrict__ x, double* __restrict__ pc) {{
    (void)x; (void)pc;
{pc_body}
  }}
{fn} atoms(const double* __restrict__ pc, const double* __restrict__ y, double* __restrict__ s) {{
    (void)pc; (void)y; (void)s;
{at_body}
  }}
{fn} tensor(const double* __restrict__ pc, const double* __restrict__ s, double* __restrict__ A) {{
    (void)pc; (void)s;
{te_body}
  }}
{fn} tensor_affine(const double* __restrict__ pc, double* __restrict__ C) {{
    (void)pc;
{af_body}
  }}
{fn} stress(const double* __restrict__ pc, const double* __restrict__ s, const double* __restrict__ e, double* __restrict__ sig) {{
{st_body}
  }}
{fn} dtheta(const double* __restrict__ x, double* __restrict__ M) {{
    (void)x;
{dt_body}
  }}
}};
"""
    return CoefficientProgram(
        dim=dim, kind=kind, stratified=dth is not None, degree=degree, ydep=ydep, natoms=natoms, npc=npc,
        ncomp=len(comps), scalar=scalar, source=src, atoms=atoms, comps=comps, dtheta=dth,
    )  # fmt: skip


# ----------------------------------------------------------------------------
# macro right-hand side f(x)  (hmm.py:131-133: L = inner(f(x), v) dx)
# ----------------------------------------------------------------------------
@dataclass
class LoadProgram:
    dim: int
    bs: int
    degree: int  # quadrature degree of inner(f, v) dx: degree(f) + 1 for the P1 test function
    source: str

    @property
    def key(self):
        return hashlib.sha1(self.source.encode()).hexdigest()[:16]


def trace_load(f, dim, bs):
    """Components of the macro right-hand side ``f(x)`` (a scalar for the Poisson classes, a ``bs``-vector for
    elasticity) as traced expressions of the macro coordinate."""
    val = ufl.call_traced(f, [("coord", "x", dim)])  # hmm.py:131: f(SpatialCoordinate(msh))
    if isinstance(val, Tensor):
        comps = [val.data[k] for k in range(val.data.shape[0])]
    elif isinstance(val, (list, tuple, np.ndarray)):
        comps = [Expr.wrap(v) for v in val]
    else:
        comps = [Expr.wrap(val)]
    if len(comps) != bs:
        raise ValueError(f"f must have {bs} component(s), got {len(comps)}")
    return comps


def build_load_program(f, dim, bs) -> LoadProgram:
    """CUDA source of ``struct HMX_LOAD`` for csrc/hmx_load_entry.cu: the device-side replacement of the FFCx kernel
    the reference assembles ``self._L`` with (hmm.py:445-450)."""
    comps = trace_load(f, dim, bs)
    degree = max(ufl.estimate_degree(c, {"x": 1}) for c in comps) + 1
    em = _Emitter({("x", k): f"x[{k}]" for k in range(3)}, "f")
    body = _body(em.lines, [em.ref(c) for c in comps], "out")
    src = f"""// hommx_b200 load program v1
struct HMX_LOAD {{
  static constexpr int DIM = {dim};
  static constexpr int BS = {bs};
  static constexpr int QDEG = {degree};
  __device__ __forceinline__ static void eval(const double* __restrict__ x, double* __restrict__ out) {{
    (void)x; (void)out;
{body}
  }}
}};
"""
    return LoadProgram(dim, bs, degree, src)
