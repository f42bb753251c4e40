"""Minimal restatement of the CasADi 3.6.4 Python API surface that the reference (awebox) uses
on its g / jac_g path.  TEST INFRASTRUCTURE (oracle) — never imported by ``awebox_b200``.

casadi==3.6.4 is a third-party dependency of the reference (``/root/reference/requirements.txt:1``,
``setup.py:51``) that is absent from this container and from /root/reference.  All of the reference's
hot-path arithmetic is *created* by calls into it (``cas.jacobian``, ``cas.mtimes``, ``cas.inv``,
``cas.Function`` ... see SURVEY.md §8c).  This package restates the published algorithms behind
those calls — dense/sparse matrix algebra over scalar expression graphs (``SX``), numeric matrices
(``DM``), ``Function`` objects with symbolic and numeric call, ``map``, forward-mode AD, structural
Jacobian sparsity with column colouring, ``collocation_points`` — so that the reference's own Python
modules can be imported from /root/reference and run here to generate golden vectors
(``oracle/make_golden.py``).  ``MX`` is an alias of ``SX`` (one graph type is enough for an oracle).
"""
import math
import itertools
import numpy as _np

from . import _sx
from ._sx import N as _N

inf = float('inf')
nan = float('nan')
pi = math.pi
__version__ = '3.6.4-shim'


class Sparsity(object):
    def __init__(self, nrow=0, ncol=0, colind=None, row=None):
        self.nrow, self.ncol = nrow, ncol
        self.colind = colind
        self.row_ = row

    @staticmethod
    def dense(nrow, ncol=1):
        return Sparsity(nrow, ncol)

    @staticmethod
    def scalar():
        return Sparsity(1, 1)

    def size1(self):
        return self.nrow

    def size2(self):
        return self.ncol

    @property
    def shape(self):
        return (self.nrow, self.ncol)


def _is_num(x):
    return isinstance(x, (int, float, _np.floating, _np.integer, bool, _np.bool_))


class _Mat(object):
    """Column-major dense matrix; elements are python floats (DM) or expression nodes (SX).
    A structural zero is the constant 0."""
    __array_priority__ = 1000.0
    __slots__ = ('n', 'm', 'e')

    # ---- construction helpers
    @classmethod
    def _make(cls, n, m, e):
        r = cls.__new__(cls)
        r.n, r.m, r.e = n, m, e
        return r

    @property
    def shape(self):
        return (self.n, self.m)

    def size(self, *a):
        if a:
            return self.n if a[0] == 1 else self.m
        return (self.n, self.m)

    def size1(self):
        return self.n

    def size2(self):
        return self.m

    def numel(self):
        return self.n * self.m

    def rows(self):
        return self.n

    def columns(self):
        return self.m

    def is_empty(self, both=False):
        return self.n * self.m == 0

    def is_scalar(self, *a):
        return self.n == 1 and self.m == 1

    def is_vector(self):
        return self.n == 1 or self.m == 1

    def is_column(self):
        return self.m == 1

    def is_dense(self):
        return True

    def __len__(self):
        return self.n if self.m == 1 or self.n > 1 else self.m

    def __deepcopy__(self, memo):
        return self._make(self.n, self.m, list(self.e))

    def __copy__(self):
        return self._make(self.n, self.m, list(self.e))

    def __iter__(self):
        raise TypeError('casadi matrices are not iterable')

    @property
    def T(self):
        n, m, e = self.n, self.m, self.e
        return self._make(m, n, [e[i + j * n] for i in range(n) for j in range(m)])

    def reshape(self, shape):
        return reshape(self, shape)

    # ---- indexing
    def _idx(self, k, dim):
        if isinstance(k, slice):
            return list(range(*k.indices(dim))), False
        if isinstance(k, (list, tuple, _np.ndarray, range)):
            return [int(i) + dim if int(i) < 0 else int(i) for i in k], False
        if isinstance(k, _Mat):
            return [int(_to_float(v)) for v in k.e], False
        k = int(k)
        if k < 0:
            k += dim
        if k < 0 or k >= dim:
            raise IndexError('index %d out of bounds for dimension %d' % (k, dim))
        return [k], True

    def __getitem__(self, key):
        if isinstance(key, tuple):
            if len(key) != 2:
                raise IndexError('bad index')
            ri, _ = self._idx(key[0], self.n)
            ci, _ = self._idx(key[1], self.m)
            n = self.n
            return self._make(len(ri), len(ci), [self.e[i + j * n] for j in ci for i in ri])
        ki, sc = self._idx(key, self.n * self.m)
        e = [self.e[k] for k in ki]
        if self.n == 1 and self.m > 1 and not sc:
            return self._make(1, len(e), e)
        return self._make(len(e), 1, e)

    def __setitem__(self, key, val):
        val = self._coerce_like(val)
        if isinstance(key, tuple):
            ri, _ = self._idx(key[0], self.n)
            ci, _ = self._idx(key[1], self.m)
            tgt = [i + j * self.n for j in ci for i in ri]
        else:
            tgt, _ = self._idx(key, self.n * self.m)
        if val.numel() == 1:
            for t in tgt:
                self.e[t] = val.e[0]
        else:
            if val.numel() != len(tgt):
                raise ValueError('setitem: size mismatch %d vs %d' % (val.numel(), len(tgt)))
            for t, v in zip(tgt, val.e):
                self.e[t] = v

    @property
    def nz(self):
        return self

    def nonzeros(self):
        return list(self.e)

    def elements(self):
        return list(self.e)

    # ---- arithmetic
    def _bin(self, op, other, swap=False):
        other = _as_mat(other)
        if other is NotImplemented:
            return NotImplemented
        a, b = (other, self) if swap else (self, other)
        return _elementwise(op, a, b)

    def __add__(self, o): return self._bin(_sx.ADD, o)
    def __radd__(self, o): return self._bin(_sx.ADD, o, True)
    def __sub__(self, o): return self._bin(_sx.SUB, o)
    def __rsub__(self, o): return self._bin(_sx.SUB, o, True)
    def __mul__(self, o): return self._bin(_sx.MUL, o)
    def __rmul__(self, o): return self._bin(_sx.MUL, o, True)
    def __truediv__(self, o): return self._bin(_sx.DIV, o)
    def __rtruediv__(self, o): return self._bin(_sx.DIV, o, True)
    __div__ = __truediv__
    __rdiv__ = __rtruediv__
    def __pow__(self, o): return self._bin(_sx.POW, o)
    def __rpow__(self, o): return self._bin(_sx.POW, o, True)
    def __neg__(self): return _unary(_sx.NEG, self)
    def __pos__(self): return self
    def __abs__(self): return _unary(_sx.FABS, self)
    def __lt__(self, o): return self._bin(_sx.LT, o)
    def __le__(self, o): return self._bin(_sx.LE, o)
    def __gt__(self, o): return self._bin(_sx.LT, o, True)
    def __ge__(self, o): return self._bin(_sx.LE, o, True)
    def __eq__(self, o):
        r = self._bin(_sx.EQ, o)
        return r
    def __ne__(self, o): return self._bin(_sx.NE, o)
    __hash__ = None

    def __matmul__(self, o): return mtimes(self, o)
    def __rmatmul__(self, o): return mtimes(o, self)

    # numpy ufunc-ish names used by ``np.sqrt(dm)`` etc.
    def sqrt(self): return _unary(_sx.SQRT, self)
    def sin(self): return _unary(_sx.SIN, self)
    def cos(self): return _unary(_sx.COS, self)
    def tan(self): return _unary(_sx.TAN, self)
    def arcsin(self): return _unary(_sx.ASIN, self)
    def arccos(self): return _unary(_sx.ACOS, self)
    def arctan(self): return _unary(_sx.ATAN, self)
    def exp(self): return _unary(_sx.EXP, self)
    def log(self): return _unary(_sx.LOG, self)
    def log10(self): return _unary(_sx.LOG, self) * (1.0 / math.log(10.0))
    def fabs(self): return _unary(_sx.FABS, self)
    def arctan2(self, o): return self._bin(_sx.ATAN2, o)


def _to_float(v):
    if isinstance(v, _N):
        if v.op != _sx.CONST:
            raise TypeError('symbolic expression cannot be converted to a number')
        return v.v
    return float(v)


class DM(_Mat):
    __slots__ = ()

    def __init__(self, *args):
        if len(args) == 0:
            self.n, self.m, self.e = 0, 0, []
            return
        if len(args) == 2 and _is_num(args[0]) and _is_num(args[1]):
            self.n, self.m = int(args[0]), int(args[1])
            self.e = [0.0] * (self.n * self.m)
            return
        x = args[0]
        if isinstance(x, Sparsity):
            self.n, self.m = x.nrow, x.ncol
            self.e = [0.0] * (self.n * self.m)
            return
        if isinstance(x, _Mat):
            self.n, self.m = x.n, x.m
            self.e = [_to_float(v) for v in x.e]
            return
        if isinstance(x, _StructViewBase):
            x = x.cat
            self.n, self.m = x.n, x.m
            self.e = [_to_float(v) for v in x.e]
            return
        if _is_num(x):
            self.n, self.m, self.e = 1, 1, [float(x)]
            return
        if isinstance(x, (list, tuple)) and any(isinstance(v, _Mat) for v in x):
            x = [float(DM(v)) if isinstance(v, _Mat) else v for v in x]
        a = _np.array(x, dtype=float)
        if a.ndim == 0:
            self.n, self.m, self.e = 1, 1, [float(a)]
        elif a.ndim == 1:
            self.n, self.m, self.e = a.shape[0], (1 if a.shape[0] > 0 else 0) or 1, [float(v) for v in a]
            if a.shape[0] == 0:
                self.n, self.m = 0, 0
        elif a.ndim == 2:
            self.n, self.m = a.shape
            self.e = [float(v) for v in a.T.reshape(-1)]
        else:
            raise ValueError('DM: too many dimensions')

    def _coerce_like(self, v):
        v = _as_mat(v)
        if isinstance(v, SX):
            raise TypeError('cannot store a symbolic expression into a DM')
        return v

    @staticmethod
    def zeros(*a):
        n, m = _shape_args(a)
        return DM._make(n, m, [0.0] * (n * m))

    @staticmethod
    def ones(*a):
        n, m = _shape_args(a)
        return DM._make(n, m, [1.0] * (n * m))

    @staticmethod
    def eye(n):
        n = int(n)
        return DM._make(n, n, [1.0 if i == j else 0.0 for j in range(n) for i in range(n)])

    @staticmethod
    def nan(*a):
        n, m = _shape_args(a)
        return DM._make(n, m, [float('nan')] * (n * m))

    @staticmethod
    def inf(*a):
        n, m = _shape_args(a)
        return DM._make(n, m, [float('inf')] * (n * m))

    def full(self):
        return _np.array(self.e, dtype=float).reshape((self.m, self.n)).T.copy()

    toarray = full

    def __array__(self, dtype=None, copy=None):
        a = self.full()
        return a if dtype is None else a.astype(dtype)

    def __float__(self):
        if self.n * self.m != 1:
            raise TypeError('only scalar DM can be converted to float (shape %s)' % (self.shape,))
        return self.e[0]

    def __int__(self):
        return int(float(self))

    __index__ = __int__

    def __bool__(self):
        if self.n * self.m != 1:
            raise TypeError('only scalar DM can be converted to bool')
        return self.e[0] != 0.0

    def is_zero(self):
        return all(v == 0.0 for v in self.e)

    def is_constant(self):
        return True

    def is_regular(self):
        return all(math.isfinite(v) for v in self.e)

    def is_eye(self):
        return self.n == self.m and all(v == (1.0 if i == j else 0.0)
                                        for j in range(self.m) for i in range(self.n)
                                        for v in [self.e[i + j * self.n]])

    def is_one(self):
        return all(v == 1.0 for v in self.e)

    def nnz(self):
        return sum(1 for v in self.e if v != 0.0) if False else self.n * self.m

    def sparsity(self):
        return Sparsity(self.n, self.m)

    def __repr__(self):
        if self.n * self.m == 1:
            return 'DM(%g)' % self.e[0]
        if self.m == 1:
            return 'DM([' + ', '.join('%g' % v for v in self.e) + '])'
        return 'DM(' + repr(self.full().tolist()) + ')'

    __str__ = __repr__


class SX(_Mat):
    __slots__ = ()

    def __init__(self, *args):
        if len(args) == 0:
            self.n, self.m, self.e = 0, 0, []
            return
        if len(args) == 2 and _is_num(args[0]) and _is_num(args[1]):
            self.n, self.m = int(args[0]), int(args[1])
            self.e = [_sx.ZERO] * (self.n * self.m)
            return
        x = args[0]
        if isinstance(x, Sparsity):
            self.n, self.m = x.nrow, x.ncol
            self.e = [_sx.ZERO] * (self.n * self.m)
            return
        if isinstance(x, SX):
            self.n, self.m, self.e = x.n, x.m, list(x.e)
            return
        if isinstance(x, _StructViewBase):
            x = _as_mat(x.cat)
            self.n, self.m, self.e = x.n, x.m, [_wrap(v) for v in x.e]
            return
        d = x if isinstance(x, DM) else DM(x)
        self.n, self.m = d.n, d.m
        self.e = [_sx.const(v) for v in d.e]

    def _coerce_like(self, v):
        v = _as_mat(v)
        if isinstance(v, DM):
            v = SX(v)
        return v

    @staticmethod
    def sym(name, *a):
        if len(a) == 0:
            n, m = 1, 1
        else:
            n, m = _shape_args(a)
        if n * m == 1:
            return SX._make(n, m, [_sx.sym(name)])
        return SX._make(n, m, [_sx.sym('%s_%d' % (name, k)) for k in range(n * m)])

    @staticmethod
    def zeros(*a):
        n, m = _shape_args(a)
        return SX._make(n, m, [_sx.ZERO] * (n * m))

    @staticmethod
    def ones(*a):
        n, m = _shape_args(a)
        return SX._make(n, m, [_sx.ONE] * (n * m))

    @staticmethod
    def eye(n):
        n = int(n)
        return SX._make(n, n, [_sx.ONE if i == j else _sx.ZERO for j in range(n) for i in range(n)])

    def is_zero(self):
        return all(_sx.is_zero(v) for v in self.e)

    def is_constant(self):
        return all(v.op == _sx.CONST for v in self.e)

    def is_symbolic(self):
        return self.n * self.m == 1 and self.e[0].op == _sx.SYM

    def is_valid_input(self):
        return all(v.op == _sx.SYM for v in self.e)

    def is_regular(self):
        return all(v.op != _sx.CONST or math.isfinite(v.v) for v in self.e)

    def nnz(self):
        return self.n * self.m

    def sparsity(self):
        return Sparsity(self.n, self.m)

    def name(self):
        return self.e[0].name

    def __float__(self):
        if self.n * self.m != 1:
            raise TypeError('only scalar SX can be converted to float')
        return _to_float(self.e[0])

    def __bool__(self):
        if self.n * self.m == 1 and self.e[0].op == _sx.CONST:
            return self.e[0].v != 0.0
        raise TypeError('symbolic SX has no truth value')

    def __array__(self, dtype=None, copy=None):
        a = _np.empty((self.n, self.m), dtype=object)
        for j in range(self.m):
            for i in range(self.n):
                a[i, j] = SX._make(1, 1, [self.e[i + j * self.n]])
        return a

    def __repr__(self):
        if self.n * self.m == 1:
            return 'SX(%s)' % _sx.to_str(self.e[0])
        if self.n * self.m <= 12:
            return 'SX(%dx%d: [%s])' % (self.n, self.m, ', '.join(_sx.to_str(v, 3) for v in self.e))
        return 'SX(%dx%d)' % (self.n, self.m)

    __str__ = __repr__


MX = SX


def _shape_args(a):
    if len(a) == 1:
        if isinstance(a[0], (tuple, list)):
            return int(a[0][0]), int(a[0][1])
        if isinstance(a[0], Sparsity):
            return a[0].nrow, a[0].ncol
        return int(a[0]), 1
    return int(a[0]), int(a[1])


def _wrap(v):
    return v if isinstance(v, _N) else _sx.const(v)


class _StructViewBase(object):
    """marker base so that matrix code can recognise casadi.tools struct views (see tools.py)."""
    pass


def _as_mat(x):
    if isinstance(x, _Mat):
        return x
    if isinstance(x, _StructViewBase):
        return _as_mat(x.cat)
    if _is_num(x):
        return DM._make(1, 1, [float(x)])
    if isinstance(x, (list, tuple)):
        if len(x) == 0:
            return DM()
        if any(isinstance(v, _Mat) for v in x):
            return vertcat(*x)
        return DM(x)
    if isinstance(x, _np.ndarray):
        if x.dtype == object:
            flat = [_as_mat(v) for v in x.reshape(-1)]
            if x.ndim <= 1:
                return vertcat(*flat)
            rows = [horzcat(*[_as_mat(v) for v in x[i, :]]) for i in range(x.shape[0])]
            return vertcat(*rows)
        return DM(x)
    return NotImplemented


def _any_sx(mats):
    return any(isinstance(m, SX) for m in mats)


def _elementwise(op, a, b):
    symbolic = isinstance(a, SX) or isinstance(b, SX)
    if a.n * a.m == 1 and b.n * b.m != 1:
        n, m = b.n, b.m
        ae = [a.e[0]] * (n * m)
        be = b.e
    elif b.n * b.m == 1 and a.n * a.m != 1:
        n, m = a.n, a.m
        ae = a.e
        be = [b.e[0]] * (n * m)
    else:
        if (a.n, a.m) != (b.n, b.m):
            if a.n * a.m == 0 and b.n * b.m == 0:
                return a
            raise ValueError('dimension mismatch: %s vs %s' % (a.shape, b.shape))
        n, m = a.n, a.m
        ae, be = a.e, b.e
    if symbolic:
        f = _sx.binary
        if not isinstance(a, SX):
            ae = [_sx.const(v) for v in ae]
        if not isinstance(b, SX):
            be = [_sx.const(v) for v in be]
        return SX._make(n, m, [f(op, x, y) for x, y in zip(ae, be)])
    f = _sx.NUM_BIN[op]
    return DM._make(n, m, [float(f(x, y)) for x, y in zip(ae, be)])


def _unary(op, a):
    a = _as_mat(a)
    if isinstance(a, SX):
        return SX._make(a.n, a.m, [_sx.unary(op, v) for v in a.e])
    f = _sx.NUM_UN[op]
    return DM._make(a.n, a.m, [float(f(v)) for v in a.e])


def _scalar_or_mat(f):
    return f


# ---- elementwise functions
def sqrt(x): return _unary(_sx.SQRT, x)
def sin(x): return _unary(_sx.SIN, x)
def cos(x): return _unary(_sx.COS, x)
def tan(x): return _unary(_sx.TAN, x)
def asin(x): return _unary(_sx.ASIN, x)
def acos(x): return _unary(_sx.ACOS, x)
def atan(x): return _unary(_sx.ATAN, x)
def exp(x): return _unary(_sx.EXP, x)
def log(x): return _unary(_sx.LOG, x)
def fabs(x): return _unary(_sx.FABS, x)
def sign(x): return _unary(_sx.SIGN, x)
def floor(x): return _unary(_sx.FLOOR, x)
def ceil(x): return _unary(_sx.CEIL, x)
def tanh(x): return _unary(_sx.TANH, x)
def sinh(x): return _unary(_sx.SINH, x)
def cosh(x): return _unary(_sx.COSH, x)
def sq(x): return _unary(_sx.SQ, x)
def log10(x): return log(x) * (1.0 / math.log(10.0))
arcsin, arccos, arctan = asin, acos, atan


def atan2(y, x): return _elementwise(_sx.ATAN2, _as_mat(y), _as_mat(x))
arctan2 = atan2
def fmin(x, y): return _elementwise(_sx.FMIN, _as_mat(x), _as_mat(y))
def fmax(x, y): return _elementwise(_sx.FMAX, _as_mat(x), _as_mat(y))
def power(x, y): return _elementwise(_sx.POW, _as_mat(x), _as_mat(y))
def times(x, y): return _elementwise(_sx.MUL, _as_mat(x), _as_mat(y))
def plus(x, y): return _elementwise(_sx.ADD, _as_mat(x), _as_mat(y))
def minus(x, y): return _elementwise(_sx.SUB, _as_mat(x), _as_mat(y))
def rdivide(x, y): return _elementwise(_sx.DIV, _as_mat(x), _as_mat(y))
def logic_and(x, y): return _elementwise(_sx.AND, _as_mat(x), _as_mat(y))
def logic_or(x, y): return _elementwise(_sx.OR, _as_mat(x), _as_mat(y))
def logic_not(x): return _unary(_sx.NOT, x)


def if_else(c, t, f, short_circuit=False):
    c, t, f = _as_mat(c), _as_mat(t), _as_mat(f)
    return _elementwise(_sx.IF_ELSE_ZERO, c, t) + _elementwise(_sx.IF_ELSE_ZERO, logic_not(c), f)


# ---- shape functions
def _cat_prepare(args):
    if len(args) == 1 and isinstance(args[0], (list, tuple)) and not _is_num(args[0]):
        args = args[0]
    mats = []
    for a in args:
        m = _as_mat(a)
        if m is NotImplemented:
            raise TypeError('cannot concatenate %r' % type(a))
        mats.append(m)
    return mats


def vertcat(*args):
    mats = [m for m in _cat_prepare(args) if not (m.n == 0 and m.m == 0)]
    if any(m.n * m.m > 0 for m in mats):
        mats = [m for m in mats if m.n > 0]
    if not mats:
        return DM()
    cls = SX if _any_sx(mats) else DM
    m = mats[0].m
    for x in mats:
        if x.m != m:
            raise ValueError('vertcat: column mismatch %s' % [y.shape for y in mats])
    n = sum(x.n for x in mats)
    e = []
    for j in range(m):
        for x in mats:
            col = x.e[j * x.n:(j + 1) * x.n]
            if cls is SX and not isinstance(x, SX):
                col = [_sx.const(v) for v in col]
            e.extend(col)
    return cls._make(n, m, e)


def horzcat(*args):
    mats = [m for m in _cat_prepare(args) if not (m.n == 0 and m.m == 0)]
    if any(m.n * m.m > 0 for m in mats):
        mats = [m for m in mats if m.m > 0]
    if not mats:
        return DM()
    cls = SX if _any_sx(mats) else DM
    n = mats[0].n
    e = []
    for x in mats:
        if x.n != n:
            raise ValueError('horzcat: row mismatch %s' % [y.shape for y in mats])
        col = x.e
        if cls is SX and not isinstance(x, SX):
            col = [_sx.const(v) for v in col]
        e.extend(col)
    return cls._make(n, sum(x.m for x in mats), e)


def veccat(*args):
    return vertcat(*[vec(m) for m in _cat_prepare(args)])


def vec(x):
    x = _as_mat(x)
    return x._make(x.n * x.m, 1, list(x.e))


def blockcat(rows):
    return vertcat(*[horzcat(*r) for r in rows])


def diagcat(*args):
    mats = _cat_prepare(args)
    n = sum(x.n for x in mats)
    m = sum(x.m for x in mats)
    cls = SX if _any_sx(mats) else DM
    out = cls.zeros(n, m)
    r = c = 0
    for x in mats:
        out[r:r + x.n, c:c + x.m] = x
        r += x.n
        c += x.m
    return out


def reshape(x, *shape):
    x = _as_mat(x)
    n, m = _shape_args(shape)
    if n == -1:
        n = x.numel() // m
    if m == -1:
        m = x.numel() // n
    if n * m != x.n * x.m:
        raise ValueError('reshape: %s -> %s' % (x.shape, (n, m)))
    return x._make(n, m, list(x.e))


def transpose(x):
    return _as_mat(x).T


def repmat(x, n, m=1):
    x = _as_mat(x)
    if isinstance(n, (tuple, list)):
        n, m = n
    col = vertcat(*([x] * int(n))) if n != 1 else x
    return horzcat(*([col] * int(m))) if m != 1 else col


def diag(x):
    x = _as_mat(x)
    if x.n == 1 or x.m == 1:
        k = x.n * x.m
        out = (SX if isinstance(x, SX) else DM).zeros(k, k)
        for i in range(k):
            out.e[i + i * k] = x.e[i]
        return out
    k = min(x.n, x.m)
    return x._make(k, 1, [x.e[i + i * x.n] for i in range(k)])


def _zero_of(x):
    return _sx.ZERO if isinstance(x, SX) else 0.0


def _is_struct_zero(v):
    return (v.op == _sx.CONST and v.v == 0.0) if isinstance(v, _N) else (v == 0.0)


def mtimes(*args):
    if len(args) == 1 and isinstance(args[0], (list, tuple)):
        args = args[0]
    out = _as_mat(args[0])
    for b in args[1:]:
        out = _mtimes2(out, _as_mat(b))
    return out


def _mtimes2(a, b):
    if a.n * a.m == 1 or b.n * b.m == 1:
        return _elementwise(_sx.MUL, a, b)
    if a.m != b.n:
        raise ValueError('mtimes: dimension mismatch %s x %s' % (a.shape, b.shape))
    symbolic = isinstance(a, SX) or isinstance(b, SX)
    n, K, m = a.n, a.m, b.m
    if not symbolic:
        r = a.full().dot(b.full())
        return DM(r)
    ae = a.e if isinstance(a, SX) else [_sx.const(v) for v in a.e]
    be = b.e if isinstance(b, SX) else [_sx.const(v) for v in b.e]
    out = []
    B, A, MUL, ADD, Z = _sx.binary, None, _sx.MUL, _sx.ADD, _sx.ZERO
    # casadi_mtimes: for each column of the result, loop over the nonzeros of that column of b,
    # then over the nonzeros of the matching column of a:  z(:,j) += a(:,k) * b(k,j)
    for j in range(m):
        colacc = [Z] * n
        for k in range(K):
            bkj = be[k + j * K]
            if _sx.is_zero(bkj):
                continue
            base = k * n
            for i in range(n):
                aik = ae[base + i]
                if _sx.is_zero(aik):
                    continue
                colacc[i] = B(ADD, colacc[i], B(MUL, aik, bkj))
        out.extend(colacc)
    return SX._make(n, m, out)


def dot(a, b):
    a, b = _as_mat(a), _as_mat(b)
    return sum1(vec(a * b))


def sum1(x):
    x = _as_mat(x)
    return mtimes(DM.ones(1, x.n), x)


def sum2(x):
    x = _as_mat(x)
    return mtimes(x, DM.ones(x.m, 1))


def sumsqr(x):
    x = _as_mat(x)
    return sum1(vec(sq(x)))


def norm_2(x):
    x = _as_mat(x)
    return sqrt(sumsqr(x))


norm_fro = norm_2


def norm_1(x):
    return sum1(vec(fabs(x)))


def norm_inf(x):
    x = _as_mat(x)
    r = fabs(x[0])
    for k in range(1, x.numel()):
        r = fmax(r, fabs(x[k]))
    return r


def cross(a, b, dim=-1):
    a, b = _as_mat(a), _as_mat(b)
    return vertcat(a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0])


def skew(a):
    a = _as_mat(a)
    return blockcat([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])


def trace(x):
    x = _as_mat(x)
    r = x[0, 0]
    for i in range(1, x.n):
        r = r + x[i, i]
    return r


def _minor(x, i, j):
    rows = [r for r in range(x.n) if r != i]
    cols = [c for c in range(x.m) if c != j]
    return det(x[rows, cols])


def det(x):
    """Matrix<Scalar>::det — minor expansion along the row/column with the fewest nonzeros."""
    x = _as_mat(x)
    n = x.n
    if n == 1:
        return x[0, 0]
    if n == 2:
        return x[0, 0] * x[1, 1] - x[0, 1] * x[1, 0]
    if isinstance(x, DM):
        return DM(float(_np.linalg.det(x.full())))
    rc = [sum(1 for j in range(n) if not _is_struct_zero(x.e[i + j * n])) for i in range(n)]
    cc = [sum(1 for i in range(n) if not _is_struct_zero(x.e[i + j * n])) for j in range(n)]
    if min(rc) == 0 or min(cc) == 0:
        return SX(0.0)
    ret = SX(0.0)
    if min(rc) <= min(cc):
        i = rc.index(min(rc))
        for j in range(n):
            if _is_struct_zero(x.e[i + j * n]):
                continue
            sgn = 1.0 if (i + j) % 2 == 0 else -1.0
            ret = ret + x[i, j] * (sgn * _minor(x, i, j))
    else:
        j = cc.index(min(cc))
        for i in range(n):
            if _is_struct_zero(x.e[i + j * n]):
                continue
            sgn = 1.0 if (i + j) % 2 == 0 else -1.0
            ret = ret + x[i, j] * (sgn * _minor(x, i, j))
    return ret


def adj(x):
    x = _as_mat(x)
    n = x.n
    out = SX.zeros(n, n)
    for i in range(n):
        for j in range(n):
            sgn = 1.0 if (i + j) % 2 == 0 else -1.0
            out[j, i] = sgn * _minor(x, i, j)
    return out


def inv_minor(x):
    return adj(x) / det(x)


def _is_tril(x):
    return all(_is_struct_zero(x.e[i + j * x.n]) for j in range(x.m) for i in range(min(j, x.n)))


def _is_triu(x):
    return all(_is_struct_zero(x.e[i + j * x.n]) for j in range(x.m) for i in range(j + 1, x.n))


def solve(a, b, *unused):
    """Matrix<Scalar>::solve: triangular substitution, else (n<=3) inverse by minor expansion,
    else QR (modified Gram-Schmidt)."""
    a, b = _as_mat(a), _as_mat(b)
    if isinstance(a, DM) and isinstance(b, DM):
        return DM(_np.linalg.solve(a.full(), b.full()))
    n = a.n
    if _is_tril(a):
        x = SX(b)
        for i in range(n):
            for k in range(b.m):
                if _is_struct_zero(x.e[i + k * n]):
                    continue
                x[i, k] = x[i, k] / a[i, i]
                for j in range(n - 1, i, -1):
                    if _is_struct_zero(a.e[j + i * n]):
                        continue
                    x[j, k] = x[j, k] - a[j, i] * x[i, k]
        return x
    if _is_triu(a):
        x = SX(b)
        for i in range(n - 1, -1, -1):
            for k in range(b.m):
                if _is_struct_zero(x.e[i + k * n]):
                    continue
                x[i, k] = x[i, k] / a[i, i]
                for j in range(0, i):
                    if _is_struct_zero(a.e[j + i * n]):
                        continue
                    x[j, k] = x[j, k] - a[j, i] * x[i, k]
        return x
    if n <= 3:
        return mtimes(inv_minor(a), b)
    q, r = qr(a)
    return solve(r, mtimes(q.T, b))


def qr(a):
    a = _as_mat(a)
    n, m = a.n, a.m
    Q = []
    R = SX.zeros(m, m)
    for i in range(m):
        ai = a[:, i]
        qi = ai
        for j in range(i):
            rij = mtimes(Q[j].T, qi)
            R[j, i] = rij
            qi = qi - rij * Q[j]
        rii = norm_2(qi)
        R[i, i] = rii
        Q.append(qi / rii)
    return horzcat(*Q), R


def inv(x):
    x = _as_mat(x)
    if isinstance(x, DM):
        return DM(_np.linalg.inv(x.full()))
    return solve(x, SX.eye(x.n))


def pinv(x):
    x = _as_mat(x)
    if isinstance(x, DM):
        return DM(_np.linalg.pinv(x.full()))
    if x.m >= x.n:
        return solve(mtimes(x, x.T), x).T
    return solve(mtimes(x.T, x), x.T)


def mpower(x, k):
    x = _as_mat(x)
    if x.numel() == 1:
        return x ** k
    r = (SX if isinstance(x, SX) else DM).eye(x.n)
    for _ in range(int(k)):
        r = mtimes(r, x)
    return r


def polyval(p, x):
    p = _as_mat(p)
    x = _as_mat(x)
    r = p[0] * 1.0
    for k in range(1, p.numel()):
        r = r * x + p[k]
    return r


def linspace(a, b, n):
    return DM(_np.linspace(float(DM(a)), float(DM(b)), int(n)))


def DM_eye(n):
    return DM.eye(n)


def SX_eye(n):
    return SX.eye(n)


MX_eye = SX_eye


# ---- symbolic queries / AD
def symvar(x):
    x = _as_mat(x)
    out = []
    seen = set()
    for n in _sx.topo([v for v in x.e]) if isinstance(x, SX) else []:
        if n.op == _sx.SYM and id(n) not in seen:
            seen.add(id(n))
            out.append(SX._make(1, 1, [n]))
    return out


def depends_on(f, x):
    f, x = _as_mat(f), _as_mat(x)
    if not isinstance(f, SX) or not isinstance(x, SX):
        return False
    xs = set(id(v) for v in x.e)
    return any(id(n) in xs for n in _sx.topo(list(f.e)))


def is_equal(a, b, depth=0):
    a, b = _as_mat(a), _as_mat(b)
    if a.shape != b.shape:
        return False
    for u, v in zip(a.e, b.e):
        if isinstance(u, _N) or isinstance(v, _N):
            if not _sx.is_equal(_wrap(u), _wrap(v), depth):
                return False
        elif u != v:
            return False
    return True


def jacobian(f, x, opts=None):
    f, x = _as_mat(f), _as_mat(x)
    nf, nx = f.n * f.m, x.n * x.m
    out = SX.zeros(nf, nx)
    if not isinstance(f, SX) or not isinstance(x, SX):
        return out
    for v in x.e:
        if v.op != _sx.SYM:
            raise ValueError('jacobian: second argument must be purely symbolic')
    rows = _sx.forward_jac(list(f.e), list(x.e))
    for i, d in enumerate(rows):
        for j, v in d.items():
            out.e[i + j * nf] = v
    return out


def gradient(f, x):
    return jacobian(f, x).T


def hessian(f, x):
    g = gradient(f, x)
    return jacobian(g, x), g


def jtimes(f, x, v, tr=False):
    J = jacobian(f, x)
    return mtimes(J.T, v) if tr else mtimes(J, v)


def substitute(ex, v, vdef):
    single = not isinstance(ex, (list, tuple))
    exs = [_as_mat(ex)] if single else [_as_mat(e) for e in ex]
    if isinstance(v, (list, tuple)):
        v = vertcat(*[vec(_as_mat(t)) for t in v])
        vdef = vertcat(*[vec(_as_mat(t)) for t in vdef])
    v, vdef = _as_mat(v), _as_mat(vdef)
    vd = SX(vdef) if not isinstance(vdef, SX) else vdef
    env = {id(s): d for s, d in zip(v.e, vd.e)}
    roots = [n for e in exs if isinstance(e, SX) for n in e.e]
    _symbolic_eval(roots, env)
    out = [SX._make(e.n, e.m, [env[id(n)] for n in e.e]) if isinstance(e, SX) else e for e in exs]
    return out[0] if single else out


def _symbolic_eval(roots, env):
    """Re-build the graph below ``roots`` with the leaves replaced per ``env`` (id -> node);
    env is completed in place for every visited node."""
    for n in _sx.topo(roots):
        k = id(n)
        if k in env:
            continue
        if n.op == _sx.CONST or n.op == _sx.SYM:
            env[k] = n
        elif n.b is None:
            a = env[id(n.a)]
            env[k] = n if a is n.a else _sx.unary(n.op, a)
        else:
            a, b = env[id(n.a)], env[id(n.b)]
            if a is n.a and b is n.b:
                env[k] = n
            elif n.op == _sx.CONSTPOW:
                env[k] = _sx.power(a, b)
            else:
                env[k] = _sx.binary(n.op, a, b)


def jacobian_sparsity(f, x):
    from . import _eval
    f, x = _as_mat(f), _as_mat(x)
    return _eval.dependency_pattern(list(f.e), list(x.e))


def collocation_points(order, scheme='radau'):
    """casadi::collocation_points — Legendre or Radau IIA points on (0,1] (integration_tools.cpp).
    casadi tabulates them as long-double literals; here they are recomputed to 40 digits and
    rounded to double."""
    import mpmath
    mpmath.mp.dps = 50
    d = int(order)
    x = mpmath.mpf
    if scheme == 'legendre':
        poly = lambda t: mpmath.legendre(d, 2 * t - 1)
    elif scheme == 'radau':
        # right Radau points: roots of P_{d-1}(s) - P_d(s), s = 2t-1  (includes s=+1)
        poly = lambda t: mpmath.legendre(d - 1, 2 * t - 1) - mpmath.legendre(d, 2 * t - 1)
    else:
        raise ValueError(scheme)
    # bracket the roots by a fine scan, then polish
    roots = []
    Ngrid = 4000
    prev_t = x(0)
    prev_v = poly(prev_t)
    for i in range(1, Ngrid + 1):
        t = x(i) / Ngrid
        v = poly(t)
        if v == 0:
            roots.append(t)
        elif prev_v * v < 0:
            roots.append(mpmath.findroot(poly, (prev_t, t), solver='anderson', tol=1e-45))
        prev_t, prev_v = t, v
    if scheme == 'radau' and (not roots or abs(roots[-1] - 1) > 1e-30):
        roots.append(x(1))
    roots = sorted(set(float(r) for r in roots))
    if len(roots) != d:
        raise RuntimeError('collocation_points: found %d roots for order %d' % (len(roots), d))
    return roots


# ---- Function
class Function(object):
    def __init__(self, name, ins, outs, *rest):
        self._name = name
        opts = {}
        in_names = out_names = None
        rest = list(rest)
        if rest and isinstance(rest[-1], dict):
            opts = rest.pop()
        if len(rest) >= 2:
            in_names, out_names = rest[0], rest[1]
        self._ins = []
        for a in ins:
            a = _as_mat(a)
            if not isinstance(a, SX):
                a = SX(a)
            for v in a.e:
                if v.op != _sx.SYM:
                    raise ValueError('Function %s: inputs must be purely symbolic' % name)
            self._ins.append(a)
        self._outs = []
        for o in outs:
            o = _as_mat(o)
            self._outs.append(o if isinstance(o, SX) else SX(o))
        self._in_names = list(in_names) if in_names else ['i%d' % k for k in range(len(self._ins))]
        self._out_names = list(out_names) if out_names else ['o%d' % k for k in range(len(self._outs))]
        self._tape = None
        self.opts = opts

    # -- metadata
    def name(self): return self._name
    def n_in(self): return len(self._ins)
    def n_out(self): return len(self._outs)
    def size_in(self, i): return self._ins[self._ii(i)].shape
    def size_out(self, i): return self._outs[self._oi(i)].shape
    def size1_in(self, i): return self._ins[self._ii(i)].n
    def size2_in(self, i): return self._ins[self._ii(i)].m
    def size1_out(self, i): return self._outs[self._oi(i)].n
    def size2_out(self, i): return self._outs[self._oi(i)].m
    def numel_in(self, i=0): return self._ins[self._ii(i)].numel()
    def numel_out(self, i=0): return self._outs[self._oi(i)].numel()
    nnz_in, nnz_out = numel_in, numel_out
    def sparsity_in(self, i): return Sparsity(*self.size_in(i))
    def sparsity_out(self, i): return Sparsity(*self.size_out(i))
    def name_in(self, i=None): return self._in_names if i is None else self._in_names[i]
    def name_out(self, i=None): return self._out_names if i is None else self._out_names[i]
    def sx_in(self, i=None): return self._ins if i is None else self._ins[self._ii(i)]
    def sx_out(self, i=None): return self._outs if i is None else self._outs[self._oi(i)]
    mx_in, mx_out = sx_in, sx_out
    def expand(self, *a): return self
    def is_null(self): return False
    def n_nodes(self): return len(_sx.topo([v for o in self._outs for v in o.e]))
    n_instructions = n_nodes

    def _ii(self, i):
        return self._in_names.index(i) if isinstance(i, str) else i

    def _oi(self, i):
        return self._out_names.index(i) if isinstance(i, str) else i

    def __repr__(self):
        return 'Function(%s: %s -> %s)' % (self._name, [a.shape for a in self._ins], [o.shape for o in self._outs])

    # -- calling
    def __call__(self, *args, **kwargs):
        if kwargs:
            args = [kwargs.get(nm, 0.0) for nm in self._in_names]
            res = self.call(list(args))
            return dict(zip(self._out_names, res))
        if len(args) == 1 and isinstance(args[0], (list, tuple)) and self.n_in() != 1:
            args = args[0]
        res = self.call(list(args))
        return res[0] if len(res) == 1 else tuple(res)

    def call(self, args, *unused):
        if len(args) != len(self._ins):
            raise ValueError('Function %s: expected %d inputs, got %d' % (self._name, len(self._ins), len(args)))
        mats = []
        for a, s in zip(args, self._ins):
            a = _as_mat(a)
            if a is NotImplemented:
                raise TypeError('Function %s: bad argument type' % self._name)
            if a.shape != s.shape:
                if a.numel() == 1:
                    a = a._make(s.n, s.m, [a.e[0]] * (s.n * s.m))
                elif a.numel() == s.numel() and (a.n == 1 or a.m == 1) and (s.n == 1 or s.m == 1):
                    a = a._make(s.n, s.m, list(a.e))
                elif a.numel() == 0:
                    a = DM.zeros(s.n, s.m)
                else:
                    raise ValueError('Function %s: argument shape %s, expected %s' % (self._name, a.shape, s.shape))
            mats.append(a)
        if _any_sx(mats):
            env = {}
            for a, s in zip(mats, self._ins):
                ae = a.e if isinstance(a, SX) else [_sx.const(v) for v in a.e]
                for sv, av in zip(s.e, ae):
                    env[id(sv)] = av
            roots = [v for o in self._outs for v in o.e]
            _symbolic_eval(roots, env)
            return [SX._make(o.n, o.m, [env[id(v)] for v in o.e]) for o in self._outs]
        from . import _eval
        if self._tape is None:
            self._tape = _eval.Tape([v for s in self._ins for v in s.e], [v for o in self._outs for v in o.e])
        x = _np.array([v for a in mats for v in a.e], dtype=float)
        y = self._tape.eval(x)
        res = []
        k = 0
        for o in self._outs:
            cnt = o.n * o.m
            res.append(DM._make(o.n, o.m, [float(v) for v in y[k:k + cnt]]))
            k += cnt
        return res

    def map(self, *args):
        # map(N), map(N, ptype), map(name, ptype, N, [], []) ...
        N = None
        for a in args:
            if isinstance(a, (int, _np.integer)) and not isinstance(a, bool):
                N = int(a)
                break
        if N is None:
            raise ValueError('Function.map: could not find N in %r' % (args,))
        return _MapFunction(self, N)

    def mapaccum(self, *a, **k):
        raise NotImplementedError('mapaccum is not part of the oracle restatement')

    def jacobian(self):
        x = vertcat(*[vec(a) for a in self._ins])
        raise NotImplementedError

    def factory(self, name, in_names, out_names, *a):
        outs = []
        for on in out_names:
            if ':' in on:
                kind, o, i = on.split(':')
                if kind != 'jac':
                    raise NotImplementedError(on)
                outs.append(jacobian(self._outs[self._oi(o)], self._ins[self._ii(i)]))
            else:
                outs.append(self._outs[self._oi(on)])
        ins = [self._ins[self._ii(i)] for i in in_names]
        return Function(name, ins, outs, in_names, out_names)

    def generate(self, *a, **k):
        raise NotImplementedError

    def stats(self):
        return {}


class _MapFunction(Function):
    """Function.map(N): horizontally concatenated inputs/outputs; evaluated column by column."""

    def __init__(self, f, N):
        self._f = f
        self._N = N
        self._name = 'map_' + f._name
        self._ins = [SX._make(a.n, a.m * N, [None] * 0) for a in f._ins]
        self._in_names = f._in_names
        self._out_names = f._out_names
        self._tape = None
        self.opts = {}

    def n_out(self): return self._f.n_out()
    def size_in(self, i): return (self._f._ins[i].n, self._f._ins[i].m * self._N)
    def size_out(self, i): return (self._f._outs[i].n, self._f._outs[i].m * self._N)

    def call(self, args, *unused):
        f, N = self._f, self._N
        mats = [_as_mat(a) for a in args]
        # CasADi evaluates a Function on arguments with npar times the expected columns npar times, side by side
        # (FunctionInternal::matching_arg / the implicit map of Function::call); awebox/pmpc.py:328-352 relies on it
        # (tracking_cost.map(N) called with N * d columns)
        npar = [a.m // (s.m * N) for a, s in zip(mats, f._ins) if a.m != s.m and a.m % (s.m * N) == 0]
        if npar and max(npar) > 1:
            N = N * max(npar)
        cols = [[] for _ in range(f.n_out())]
        for k in range(N):
            sub = []
            for a, s in zip(mats, f._ins):
                if a.m == s.m * N:
                    sub.append(a[:, k * s.m:(k + 1) * s.m])
                elif a.m == s.m:      # non-repeated input is broadcast
                    sub.append(a)
                else:
                    raise ValueError('map: argument shape %s vs %s x %d' % (a.shape, s.shape, N))
            res = f.call(sub)
            for i, r in enumerate(res):
                cols[i].append(r)
        return [horzcat(*c) for c in cols]


def external(*a, **k):
    raise NotImplementedError('casadi.external is not part of the oracle restatement')


def nlpsol(*a, **k):
    raise NotImplementedError('casadi.nlpsol (IPOPT) is not part of the oracle restatement')


def rootfinder(*a, **k):
    raise NotImplementedError('casadi.rootfinder is not part of the oracle restatement')


def integrator(*a, **k):
    raise NotImplementedError('casadi.integrator is not part of the oracle restatement')


def interpolant(*a, **k):
    raise NotImplementedError('casadi.interpolant is not part of the oracle restatement')


def nlpsol_out(*a):
    return ['x', 'f', 'g', 'lam_x', 'lam_g', 'lam_p']


def nlpsol_n_out():
    return 6


class Callback(Function):
    def __init__(self):
        pass

    def construct(self, name, opts=None):
        self._name = name


class GlobalOptions(object):
    @staticmethod
    def setSimplificationOnTheFly(flag):
        pass
