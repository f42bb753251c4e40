"""Scalar expression graph used to derive the node model and to emit the CUDA / C code.

This is the product's own derivation engine (it replaces what the reference delegates to
CasADi's SX type: ``cas.jacobian`` at awebox/mdl/lagr_dyn_dir/lagr_dyn.py:72-75,
awebox/mdl/lagr_dyn_dir/tools.py:41,70, awebox/tools/vector_operations.py:260).  Design:

* every node is hash-consed, so common sub-expressions are shared by construction and the
  emitted straight-line code is already CSE'd;
* vectors / matrices are numpy object arrays of ``Expr`` (numpy does matmul, slicing, reshape);
* ``jacobian`` is sparse forward-mode AD over the DAG;
* ``dependency`` gives the *structural* sparsity (which inputs an output can reach), which is the
  definition CasADi uses for the pattern of ``jac_g``.

Only host-side, build-time code: the hot path runs the emitted kernels.
"""
import math
import numpy as np

# op codes
CONST, SYM, ADD, SUB, MUL, DIV, NEG, SQRT, POWC, POW, SIN, COS, TAN, ASIN, ACOS, ATAN, ATAN2, EXP, LOG, FABS, \
    SIGN, RSQRT = range(22)

_UNARY = {NEG, SQRT, SIN, COS, TAN, ASIN, ACOS, ATAN, EXP, LOG, FABS, SIGN, RSQRT}
_NAMES = {ADD: '+', SUB: '-', MUL: '*', DIV: '/', NEG: 'neg', SQRT: 'sqrt', POWC: 'pow', POW: 'pow', SIN: 'sin',
          COS: 'cos', TAN: 'tan', ASIN: 'asin', ACOS: 'acos', ATAN: 'atan', ATAN2: 'atan2', EXP: 'exp',
          LOG: 'log', FABS: 'fabs', SIGN: 'sign', RSQRT: 'rsqrt'}

_table = {}
_count = [0]
# scheduling region of newly created nodes (see ``region``); 0 = no region
_region = [0]
_region_names = ['']


class Expr(object):
    __slots__ = ('op', 'a', 'b', 'val', 'name', 'idx', 'reg', '__weakref__')
    __array_priority__ = 100.0

    def __init__(self, op, a=None, b=None, val=None, name=None):
        self.op, self.a, self.b, self.val, self.name = op, a, b, val, name
        self.idx = _count[0]
        self.reg = _region[0]
        _count[0] += 1

    # ---- arithmetic (numpy object arrays broadcast elementwise; numpy scalars defer to us)
    __array_ufunc__ = None

    def __add__(self, o): return _bc(o, lambda v: add(self, _e(v)))
    def __radd__(self, o): return _bc(o, lambda v: add(_e(v), self))
    def __sub__(self, o): return _bc(o, lambda v: sub(self, _e(v)))
    def __rsub__(self, o): return _bc(o, lambda v: sub(_e(v), self))
    def __mul__(self, o): return _bc(o, lambda v: mul(self, _e(v)))
    def __rmul__(self, o): return _bc(o, lambda v: mul(_e(v), self))
    def __truediv__(self, o): return _bc(o, lambda v: div(self, _e(v)))
    def __rtruediv__(self, o): return _bc(o, lambda v: div(_e(v), self))
    def __neg__(self): return neg(self)
    def __pos__(self): return self
    def __pow__(self, o): return power(self, _e(o))
    def __rpow__(self, o): return power(_e(o), self)

    def is_const(self): return self.op == CONST
    def is_zero(self): return self.op == CONST and self.val == 0.0
    def is_one(self): return self.op == CONST and self.val == 1.0

    def __repr__(self):
        return to_str(self, 5)

    def __float__(self):
        if self.op != CONST:
            raise TypeError('not a constant expression')
        return self.val

    def __hash__(self):
        return self.idx

    def __eq__(self, o):
        return self is o

    def __deepcopy__(self, memo):
        return self


class region(object):
    """``with region('tether_elem_3'):`` — every node first created inside the block is tagged with a
    new region id.  Regions are a hint for the code generator's scheduler only (nodes of one region are
    emitted together, so the values private to the region die inside it); ``jacobian`` gives the
    derivative nodes of a tagged node the same tag.  Nested blocks: innermost wins."""

    def __init__(self, name):
        _region_names.append(name)
        self.rid = len(_region_names) - 1

    def __enter__(self):
        self.prev = _region[0]
        _region[0] = self.rid
        return self

    def __exit__(self, *a):
        _region[0] = self.prev
        return False


def region_name(rid):
    return _region_names[rid]


def _bc(o, f):
    if isinstance(o, np.ndarray):
        out = np.empty(o.shape, dtype=object)
        flat_in, flat_out = o.reshape(-1), out.reshape(-1)
        for i in range(flat_in.shape[0]):
            flat_out[i] = f(flat_in[i])
        return out
    return f(o)


def _e(x):
    if isinstance(x, Expr):
        return x
    if isinstance(x, (int, float, np.floating, np.integer)):
        return const(x)
    raise TypeError('cannot convert %r to Expr' % (type(x),))


def const(v):
    v = float(v)
    key = (CONST, repr(v))
    n = _table.get(key)
    if n is None:
        n = Expr(CONST, val=v)
        _table[key] = n
    return n


ZERO = const(0.0)
ONE = const(1.0)
MONE = const(-1.0)
TWO = const(2.0)


def sym(name):
    return Expr(SYM, name=name)


def symvec(name, n):
    return np.array([sym('%s_%d' % (name, i)) for i in range(n)], dtype=object)


def _mk(op, a, b=None):
    key = (op, a.idx, b.idx if b is not None else -1)
    n = _table.get(key)
    if n is None:
        n = Expr(op, a, b)
        _table[key] = n
    return n


def add(a, b):
    if a.op == CONST and b.op == CONST:
        return const(a.val + b.val)
    if a.is_zero():
        return b
    if b.is_zero():
        return a
    if b.op == NEG:
        return sub(a, b.a)
    if a.op == NEG:
        return sub(b, a.a)
    # canonical operand order: a constant first, otherwise by creation index (constants are cached across
    # models, so their index must not decide: the generated text would depend on what was built before)
    if b.op == CONST or (a.op != CONST and a.idx > b.idx):
        a, b = b, a
    return _mk(ADD, a, b)


def sub(a, b):
    if a.op == CONST and b.op == CONST:
        return const(a.val - b.val)
    if b.is_zero():
        return a
    if a.is_zero():
        return neg(b)
    if a is b:
        return ZERO
    if b.op == NEG:
        return add(a, b.a)
    if a.op == ADD and a.b is b:
        return a.a
    if a.op == ADD and a.a is b:
        return a.b
    return _mk(SUB, a, b)


def neg(a):
    if a.op == CONST:
        return const(-a.val)
    if a.op == NEG:
        return a.a
    if a.op == SUB:
        return sub(a.b, a.a)
    return _mk(NEG, a)


def mul(a, b):
    if a.op == CONST and b.op == CONST:
        return const(a.val * b.val)
    if a.is_zero() or b.is_zero():
        return ZERO
    if a.is_one():
        return b
    if b.is_one():
        return a
    if a.op == CONST and a.val == -1.0:
        return neg(b)
    if b.op == CONST and b.val == -1.0:
        return neg(a)
    if a.op == NEG and b.op == NEG:
        return mul(a.a, b.a)
    if a.op == NEG:
        return neg(mul(a.a, b))
    if b.op == NEG:
        return neg(mul(a, b.a))
    # constant folding through products  c1*(c2*x) -> (c1*c2)*x
    if a.op == CONST and b.op == MUL and b.a.op == CONST:
        return mul(const(a.val * b.a.val), b.b)
    if b.op == CONST and a.op == MUL and a.a.op == CONST:
        return mul(const(b.val * a.a.val), a.b)
    if b.op == CONST or (a.op != CONST and a.idx > b.idx):
        a, b = b, a
    return _mk(MUL, a, b)


def div(a, b):
    if b.op == CONST:
        if b.val == 0.0:
            return const(float('nan'))
        if a.op == CONST:
            return const(a.val / b.val)
        return mul(const(1.0 / b.val), a)        # FP64 division is expensive on the GPU
    if a.is_zero():
        return ZERO
    if a is b:
        return ONE
    if a.op == NEG:
        return neg(div(a.a, b))
    if b.op == NEG:
        return neg(div(a, b.a))
    if b.op == SQRT:
        return mul(a, _mk(RSQRT, b.a))
    return _mk(DIV, a, b)


def sqrt(a):
    a = _e(a)
    if a.op == CONST:
        return const(math.sqrt(a.val)) if a.val >= 0 else const(float('nan'))
    return _mk(SQRT, a)


def rsqrt(a):
    a = _e(a)
    if a.op == CONST:
        return const(1.0 / math.sqrt(a.val))
    return _mk(RSQRT, a)


def power(a, b):
    a, b = _e(a), _e(b)
    if b.op == CONST:
        p = b.val
        if a.op == CONST:
            return const(math.pow(a.val, p))
        if p == 0.0:
            return ONE
        if p == 1.0:
            return a
        if p == 0.5:
            return sqrt(a)
        if p == -0.5:
            return rsqrt(a)
        if p == 1.5:
            return mul(a, sqrt(a))
        if p == int(p) and abs(p) <= 64:
            n = int(abs(p))
            result = None
            base = a
            while n:
                if n & 1:
                    result = base if result is None else mul(result, base)
                n >>= 1
                if n:
                    base = mul(base, base)
            return result if p > 0 else div(ONE, result)
        return _mk(POWC, a, b)
    return _mk(POW, a, b)


def _un(op, f):
    def g(a):
        a = _e(a)
        if a.op == CONST:
            return const(f(a.val))
        return _mk(op, a)
    return g


sin = _un(SIN, math.sin)
cos = _un(COS, math.cos)
tan = _un(TAN, math.tan)
asin = _un(ASIN, math.asin)
acos = _un(ACOS, math.acos)
atan = _un(ATAN, math.atan)
exp = _un(EXP, math.exp)
log = _un(LOG, math.log)
fabs = _un(FABS, abs)
sign = _un(SIGN, lambda v: (v > 0) - (v < 0))


def atan2(a, b):
    a, b = _e(a), _e(b)
    if a.op == CONST and b.op == CONST:
        return const(math.atan2(a.val, b.val))
    return _mk(ATAN2, a, b)


def to_str(n, depth=4):
    if n.op == CONST:
        return repr(n.val)
    if n.op == SYM:
        return n.name
    if depth <= 0:
        return '@%d' % n.idx
    if n.b is not None:
        if n.op in (ADD, SUB, MUL, DIV):
            return '(%s%s%s)' % (to_str(n.a, depth - 1), _NAMES[n.op], to_str(n.b, depth - 1))
        return '%s(%s,%s)' % (_NAMES[n.op], to_str(n.a, depth - 1), to_str(n.b, depth - 1))
    return '%s(%s)' % (_NAMES[n.op], to_str(n.a, depth - 1))


# ------------------------------------------------------------------ graph utilities
def topo(outputs):
    """dependencies-first ordering of all nodes reachable from ``outputs`` (depth-first, in the
    order the outputs are given — this is also the emission order of the code generator)."""
    order, seen = [], set()
    for root in outputs:
        if root.idx in seen:
            continue
        stack = [(root, False)]
        while stack:
            n, done = stack.pop()
            if done:
                order.append(n)
                continue
            if n.idx in seen:
                continue
            seen.add(n.idx)
            stack.append((n, True))
            if n.b is not None and n.b.idx not in seen:
                stack.append((n.b, False))
            if n.a is not None and n.a.idx not in seen:
                stack.append((n.a, False))
    return order


def partials(n):
    op, a, b = n.op, n.a, n.b
    if op == ADD: return ONE, ONE
    if op == SUB: return ONE, MONE
    if op == MUL: return b, a
    if op == DIV: return div(ONE, b), neg(div(n, b))
    if op == NEG: return MONE, None
    if op == SQRT: return div(const(0.5), n), None
    if op == RSQRT: return mul(const(-0.5), div(n, a)), None
    if op == POWC: return mul(b, div(n, a)), ZERO
    if op == POW: return mul(b, div(n, a)), mul(log(a), n)       # b a^(b-1) as b a^b / a: reuses the value
    if op == SIN: return cos(a), None
    if op == COS: return neg(sin(a)), None
    if op == TAN: return add(ONE, mul(n, n)), None
    if op == ASIN: return rsqrt(sub(ONE, mul(a, a))), None
    if op == ACOS: return neg(rsqrt(sub(ONE, mul(a, a)))), None
    if op == ATAN: return div(ONE, add(ONE, mul(a, a))), None
    if op == ATAN2:
        t = add(mul(a, a), mul(b, b))
        return div(b, t), neg(div(a, t))
    if op == EXP: return n, None
    if op == LOG: return div(ONE, a), None
    if op == FABS: return sign(a), None
    if op == SIGN: return ZERO, None
    raise NotImplementedError(op)


def jacobian(exprs, xs):
    """Sparse forward-mode Jacobian. ``exprs``: iterable of Expr; ``xs``: iterable of SYM Expr.
    Returns a list (per expr) of dicts {column: Expr} holding the symbolically non-zero entries."""
    exprs = [_e(v) for v in np.asarray(exprs, dtype=object).reshape(-1)]
    xs = list(np.asarray(xs, dtype=object).reshape(-1))
    col = {x.idx: j for j, x in enumerate(xs)}
    D = {}
    for n in topo(exprs):
        if n.op == SYM:
            j = col.get(n.idx)
            if j is not None:
                D[n.idx] = {j: ONE}
            continue
        if n.op == CONST:
            continue
        da = D.get(n.a.idx)
        db = D.get(n.b.idx) if n.b is not None else None
        if not da and not db:
            continue
        saved, _region[0] = _region[0], n.reg
        pa, pb = partials(n)
        d = {}
        if da and not pa.is_zero():
            for j, v in da.items():
                d[j] = mul(pa, v)
        if db and pb is not None and not pb.is_zero():
            for j, v in db.items():
                t = mul(pb, v)
                d[j] = add(d[j], t) if j in d else t
        _region[0] = saved
        d = {j: v for j, v in d.items() if not v.is_zero()}
        if d:
            D[n.idx] = d
    return [D.get(e.idx, {}) for e in exprs]


def gradient(expr, xs):
    """Reverse-mode gradient of the scalar ``expr`` with respect to the SYM nodes ``xs``: one sweep over the DAG in
    reverse topological order.  Returns {column: Expr} with the symbolically non-zero entries."""
    expr = _e(expr)
    xs = list(np.asarray(xs, dtype=object).reshape(-1))
    bar = {expr.idx: ONE}
    for n in reversed(topo([expr])):
        bn = bar.get(n.idx)
        if bn is None or n.op in (CONST, SYM) or bn.is_zero():
            continue
        saved, _region[0] = _region[0], n.reg
        pa, pb = partials(n)
        for dep, pd in ((n.a, pa), (n.b, pb)):
            if dep is None or pd is None or dep.op == CONST or pd.is_zero():
                continue
            t = mul(pd, bn)
            old = bar.get(dep.idx)
            bar[dep.idx] = t if old is None else add(old, t)
        _region[0] = saved
    out = {}
    for j, x in enumerate(xs):
        v = bar.get(x.idx)
        if v is not None and not v.is_zero():
            out[j] = v
    return out


def hessian_triu(expr, xs):
    """Upper triangle of the Hessian of the scalar ``expr``: forward-mode Jacobian of the reverse-mode gradient.
    Returns {(a, b): Expr} with a <= b (the union of the symbolic entries of both triangles)."""
    g = gradient(expr, xs)
    cols = sorted(g)
    rows = jacobian([g[c] for c in cols], xs)
    out = {}
    for a, d in zip(cols, rows):
        for b, v in d.items():
            key = (a, b) if a <= b else (b, a)
            if key not in out or a <= b:
                out[key] = v
    return out


def jacobian_dense(exprs, xs):
    exprs = list(np.asarray(exprs, dtype=object).reshape(-1))
    xs = list(np.asarray(xs, dtype=object).reshape(-1))
    rows = jacobian(exprs, xs)
    J = np.empty((len(exprs), len(xs)), dtype=object)
    J[:, :] = ZERO
    for i, d in enumerate(rows):
        for j, v in d.items():
            J[i, j] = v
    return J


def dependency(exprs, xs):
    """Structural dependency: list (per expr) of sorted column indices of ``xs`` reachable from it."""
    exprs = [_e(v) for v in np.asarray(exprs, dtype=object).reshape(-1)]
    xs = list(np.asarray(xs, dtype=object).reshape(-1))
    col = {x.idx: j for j, x in enumerate(xs)}
    dep = {}
    for n in topo(exprs):
        if n.op == SYM:
            j = col.get(n.idx)
            dep[n.idx] = (1 << j) if j is not None else 0
        elif n.op == CONST:
            dep[n.idx] = 0
        elif n.b is None:
            dep[n.idx] = dep[n.a.idx]
        else:
            dep[n.idx] = dep[n.a.idx] | dep[n.b.idx]
    out = []
    for e in exprs:
        bits = dep[e.idx]
        cols = []
        while bits:
            low = bits & -bits
            cols.append(low.bit_length() - 1)
            bits ^= low
        out.append(cols)
    return out


def substitute(exprs, mapping):
    """Rebuild ``exprs`` with SYM leaves replaced according to ``mapping`` {sym Expr: Expr}."""
    flat = [_e(v) for v in np.asarray(exprs, dtype=object).reshape(-1)]
    env = {k.idx: v for k, v in mapping.items()}
    for n in topo(flat):
        if n.idx in env:
            continue
        if n.op in (CONST, SYM):
            env[n.idx] = n
            continue
        saved, _region[0] = _region[0], n.reg
        if n.b is None:
            env[n.idx] = _rebuild(n.op, env[n.a.idx], None)
        else:
            env[n.idx] = _rebuild(n.op, env[n.a.idx], env[n.b.idx])
        _region[0] = saved
    return [env[e.idx] for e in flat]


_BUILD = {ADD: add, SUB: sub, MUL: mul, DIV: div, POWC: power, POW: power, ATAN2: atan2}
_BUILD1 = {NEG: neg, SQRT: sqrt, RSQRT: rsqrt, SIN: sin, COS: cos, TAN: tan, ASIN: asin, ACOS: acos, ATAN: atan,
           EXP: exp, LOG: log, FABS: fabs, SIGN: sign}


def _rebuild(op, a, b):
    if b is None:
        return _BUILD1[op](a)
    return _BUILD[op](a, b)


# ------------------------------------------------------------------ numeric evaluation (host, tests)
_NP1 = {NEG: np.negative, SQRT: np.sqrt, RSQRT: lambda x: 1.0 / np.sqrt(x), SIN: np.sin, COS: np.cos, TAN: np.tan,
        ASIN: np.arcsin, ACOS: np.arccos, ATAN: np.arctan, EXP: np.exp, LOG: np.log, FABS: np.abs, SIGN: np.sign}
_NP2 = {ADD: np.add, SUB: np.subtract, MUL: np.multiply, DIV: np.divide, POWC: np.power, POW: np.power,
        ATAN2: np.arctan2}


def evaluate(exprs, xs, values):
    """Evaluate with numpy; ``values`` is [len(xs)] or [len(xs), B]."""
    flat = [_e(v) for v in np.asarray(exprs, dtype=object).reshape(-1)]
    xs = list(np.asarray(xs, dtype=object).reshape(-1))
    values = np.asarray(values, dtype=float)
    env = {x.idx: values[j] for j, x in enumerate(xs)}
    with np.errstate(all='ignore'):
        for n in topo(flat):
            if n.idx in env:
                continue
            if n.op == CONST:
                env[n.idx] = n.val
            elif n.op == SYM:
                raise ValueError('free symbol %s' % n.name)
            elif n.b is None:
                env[n.idx] = _NP1[n.op](env[n.a.idx])
            else:
                env[n.idx] = _NP2[n.op](env[n.a.idx], env[n.b.idx])
    shape = values.shape[1:]
    return np.stack([np.broadcast_to(np.asarray(env[e.idx], dtype=float), shape) for e in flat])


def count_ops(exprs):
    """Operation census of the CSE'd graph below ``exprs`` (flops: add/sub/mul/neg = 1, div/sqrt/
    rsqrt = 1 each as *listed* operations, transcendental = 1)."""
    census = {}
    for n in topo([_e(v) for v in np.asarray(exprs, dtype=object).reshape(-1)]):
        if n.op in (CONST, SYM):
            continue
        census[_NAMES[n.op]] = census.get(_NAMES[n.op], 0) + 1
    return census


# ------------------------------------------------------------------ small vector helpers (numpy object arrays)
def vec(items):
    return np.array([_e(v) for v in items], dtype=object)


def zeros(*shape):
    a = np.empty(shape, dtype=object)
    a[...] = ZERO
    return a


def cvec(values):
    return np.array([const(v) for v in np.asarray(values, dtype=float).reshape(-1)], dtype=object)


def dot(a, b):
    r = ZERO
    for x, y in zip(a, b):
        r = r + x * y
    return r


def cross(a, b):
    return vec([a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]])


def matvec(M, v):
    return vec([dot(M[i, :], v) for i in range(M.shape[0])])


def matmul(A, B):
    out = zeros(A.shape[0], B.shape[1])
    for i in range(A.shape[0]):
        for j in range(B.shape[1]):
            out[i, j] = dot(A[i, :], B[:, j])
    return out


def skew(v):
    return np.array([[ZERO, -v[2], v[1]], [v[2], ZERO, -v[0]], [-v[1], v[0], ZERO]], dtype=object)
