"""Scalar symbolic core of the CasADi restatement (TEST INFRASTRUCTURE — oracle only).

This file restates, from the published behaviour of casadi==3.6.4 (``casadi/core/sx_elem.cpp``,
``casadi/core/calculus.hpp``; source NOT present under /root/reference, pinned by
``/root/reference/requirements.txt:1``), the scalar expression type ``SXElem``:

* expression nodes with the elementary operation set the reference uses
  (+ - * / neg sq sqrt pow sin cos tan asin acos atan atan2 exp log fabs fmin fmax sign ...),
* the construction-time simplifications (``x+0``, ``0*x``, ``x-x`` at equality depth 1,
  constant folding, integer powers by repeated squaring, ...),
* the partial-derivative table used by algorithmic differentiation.

Nothing in the shipped product (``awebox_b200/``) imports this module.
"""
import math

# op codes
CONST, SYM = 0, 1
ADD, SUB, MUL, DIV, POW, ATAN2, FMIN, FMAX, LT, LE, EQ, NE, AND, OR, CONSTPOW = range(10, 25)
NEG, SQ, SQRT, SIN, COS, TAN, ASIN, ACOS, ATAN, EXP, LOG, FABS, SIGN, INV, NOT, FLOOR, CEIL, \
    TANH, SINH, COSH, IF_ELSE_ZERO = range(40, 61)

BINARY = {ADD, SUB, MUL, DIV, POW, ATAN2, FMIN, FMAX, LT, LE, EQ, NE, AND, OR, CONSTPOW, IF_ELSE_ZERO}
COMMUTATIVE = {ADD, MUL, FMIN, FMAX, EQ, NE, AND, OR}

OPNAME = {ADD: '+', SUB: '-', MUL: '*', DIV: '/', POW: 'pow', ATAN2: 'atan2', FMIN: 'fmin', FMAX: 'fmax',
          LT: '<', LE: '<=', EQ: '==', NE: '!=', AND: '&&', OR: '||', CONSTPOW: 'pow',
          NEG: '-', SQ: 'sq', SQRT: 'sqrt', SIN: 'sin', COS: 'cos', TAN: 'tan', ASIN: 'asin', ACOS: 'acos',
          ATAN: 'atan', EXP: 'exp', LOG: 'log', FABS: 'fabs', SIGN: 'sign', INV: 'inv', NOT: '!',
          FLOOR: 'floor', CEIL: 'ceil', TANH: 'tanh', SINH: 'sinh', COSH: 'cosh', IF_ELSE_ZERO: 'if_else_zero'}

_counter = [0]


class N(object):
    """One SXElem node. ``op`` selects the operation, ``a``/``b`` are the dependencies,
    ``v`` is the value of a constant, ``name`` the name of a symbol."""
    __slots__ = ('op', 'a', 'b', 'v', 'name', 'id')

    def __init__(self, op, a=None, b=None, v=None, name=None):
        self.op = op
        self.a = a
        self.b = b
        self.v = v
        self.name = name
        _counter[0] += 1
        self.id = _counter[0]

    def __deepcopy__(self, memo):
        return self

    def __copy__(self):
        return self

    def __repr__(self):
        return to_str(self, 6)

    # constants
    def is_const(self):
        return self.op == CONST


_const_cache = {}


def const(v):
    """ConstantSX nodes are cached by value (sx_elem.cpp: IntegerSX / RealtypeSX caches)."""
    v = float(v)
    key = v if v == v else 'nan'
    if key == 0.0:
        key = 0.0 if math.copysign(1.0, v) > 0 else '-0'
    n = _const_cache.get(key)
    if n is None:
        n = N(CONST, v=v)
        _const_cache[key] = n
    return n


ZERO = const(0.0)
ONE = const(1.0)
MINUS_ONE = const(-1.0)
TWO = const(2.0)
HALF = const(0.5)
NAN = const(float('nan'))


def sym(name):
    return N(SYM, name=name)


def is_zero(x):
    return x.op == CONST and x.v == 0.0


def is_one(x):
    return x.op == CONST and x.v == 1.0


def is_minus_one(x):
    return x.op == CONST and x.v == -1.0


def is_op(x, op):
    return x.op == op


def is_doubled(x):
    return x.op == ADD and is_equal(x.a, x.b, 1)


def is_equal(x, y, depth=1):
    """SXElem::is_equal with SXNode::eq_depth_ == 1."""
    if x is y:
        return True
    if depth <= 0:
        return False
    if x.op != y.op:
        return False
    if x.op == CONST:
        return x.v == y.v
    if x.op == SYM:
        return False
    if x.op in BINARY:
        if is_equal(x.a, y.a, depth - 1) and is_equal(x.b, y.b, depth - 1):
            return True
        if x.op in COMMUTATIVE and is_equal(x.a, y.b, depth - 1) and is_equal(x.b, y.a, depth - 1):
            return True
        return False
    return is_equal(x.a, y.a, depth - 1)


# ---------------------------------------------------------------- numeric kernels
def _sign(x):
    return 1.0 if x > 0 else (-1.0 if x < 0 else (0.0 if x == 0 else float('nan')))


def _safe_div(x, y):
    try:
        return x / y
    except ZeroDivisionError:
        if x == 0 or x != x:
            return float('nan')
        return math.copysign(float('inf'), x) * math.copysign(1.0, y)


def _safe_pow(x, y):
    try:
        return math.pow(x, y)
    except (ValueError, OverflowError, ZeroDivisionError):
        return float('nan')


def _safe(f):
    def g(x):
        try:
            return f(x)
        except (ValueError, OverflowError):
            return float('nan')
    return g


NUM_BIN = {
    ADD: lambda x, y: x + y, SUB: lambda x, y: x - y, MUL: lambda x, y: x * y, DIV: _safe_div,
    POW: _safe_pow, CONSTPOW: _safe_pow, ATAN2: math.atan2, FMIN: min, FMAX: max,
    LT: lambda x, y: float(x < y), LE: lambda x, y: float(x <= y), EQ: lambda x, y: float(x == y),
    NE: lambda x, y: float(x != y), AND: lambda x, y: float(bool(x) and bool(y)),
    OR: lambda x, y: float(bool(x) or bool(y)), IF_ELSE_ZERO: lambda x, y: y if x != 0 else 0.0,
}
NUM_UN = {
    NEG: lambda x: -x, SQ: lambda x: x * x, SQRT: _safe(math.sqrt), SIN: math.sin, COS: math.cos,
    TAN: math.tan, ASIN: _safe(math.asin), ACOS: _safe(math.acos), ATAN: math.atan, EXP: _safe(math.exp),
    LOG: _safe(lambda x: math.log(x) if x > 0 else (float('-inf') if x == 0 else float('nan'))),
    FABS: abs, SIGN: _sign, INV: lambda x: _safe_div(1.0, x), NOT: lambda x: float(not x),
    FLOOR: math.floor, CEIL: math.ceil, TANH: math.tanh, SINH: _safe(math.sinh), COSH: _safe(math.cosh),
}


# ---------------------------------------------------------------- node construction
def unary(op, x):
    if x.op == CONST:
        return const(NUM_UN[op](x.v))
    if op == NEG:
        if x.op == NEG:
            return x.a
        if x.op == SUB:                 # -(a-b) is kept as a node in casadi; keep
            pass
    elif op == SQ:
        if x.op == SQRT:
            return x.a
        if x.op == NEG:
            return unary(SQ, x.a)
    elif op == SQRT:
        if x.op == SQ:
            return unary(FABS, x.a)
    elif op == FABS:
        if x.op == FABS or x.op == SQ:
            return x
    elif op == INV:
        if x.op == INV:
            return x.a
    return N(op, x)


def binary(op, x, y):
    """SXElem::binary with GlobalOptions::simplification_on_the_fly == true."""
    if x.op == CONST and y.op == CONST:
        return const(NUM_BIN[op](x.v, y.v))
    if op == ADD:
        if is_zero(x):
            return y
        if is_zero(y):
            return x
        if y.op == NEG:
            return binary(SUB, x, y.a)
        if x.op == NEG:
            return binary(SUB, y, x.a)
        if x.op == MUL and y.op == MUL and x.a.op == CONST and x.a.v == 0.5 and y.a.op == CONST \
                and y.a.v == 0.5 and is_equal(y.b, x.b, 1):
            return x.b
        if x.op == DIV and y.op == DIV and x.b.op == CONST and x.b.v == 2 and y.b.op == CONST \
                and y.b.v == 2 and is_equal(y.a, x.a, 1):
            return x.a
        if x.op == SUB and is_equal(x.b, y, 1):
            return x.a
        if y.op == SUB and is_equal(x, y.b, 1):
            return y.a
        if x.op == SQ and y.op == SQ and ((x.a.op == SIN and y.a.op == COS) or (x.a.op == COS and y.a.op == SIN)) \
                and is_equal(x.a.a, y.a.a, 1):
            return ONE
    elif op == SUB:
        if is_zero(y):
            return x
        if is_zero(x):
            return unary(NEG, y)
        if is_equal(x, y, 1):
            return ZERO
        if y.op == NEG:
            return binary(ADD, x, y.a)
        if x.op == ADD and is_equal(x.b, y, 1):
            return x.a
        if x.op == ADD and is_equal(x.a, y, 1):
            return x.b
        if y.op == ADD and is_equal(x, y.b, 1):
            return unary(NEG, y.a)
        if y.op == ADD and is_equal(x, y.a, 1):
            return unary(NEG, y.b)
        if x.op == NEG:
            return unary(NEG, binary(ADD, x.a, y))
    elif op == MUL:
        if is_equal(y, x, 1):
            return unary(SQ, x)
        if x.op != CONST and y.op == CONST:
            return binary(MUL, y, x)
        if is_zero(x) or is_zero(y):
            return ZERO
        if is_one(x):
            return y
        if is_one(y):
            return x
        if is_minus_one(y):
            return unary(NEG, x)
        if is_minus_one(x):
            return unary(NEG, y)
        if y.op == INV:
            return binary(DIV, x, y.a)
        if x.op == INV:
            return binary(DIV, y, x.a)
        if x.op == CONST and y.op == MUL and y.a.op == CONST and x.v * y.a.v == 1:
            return y.b
        if x.op == CONST and y.op == DIV and y.b.op == CONST and x.v == y.b.v:
            return y.a
        if x.op == DIV and is_equal(x.b, y, 1):
            return x.a
        if y.op == DIV and is_equal(y.b, x, 1):
            return y.a
        if x.op == NEG:
            return unary(NEG, binary(MUL, x.a, y))
        if y.op == NEG:
            return unary(NEG, binary(MUL, x, y.a))
    elif op == DIV:
        if is_zero(y):
            return NAN
        if is_zero(x):
            return ZERO
        if is_one(y):
            return x
        if is_minus_one(y):
            return unary(NEG, x)
        if is_equal(x, y, 1):
            return ONE
        if is_doubled(x) and y.op == CONST and y.v == 2:
            return x.a
        if x.op == MUL and is_equal(y, x.a, 1):
            return x.b
        if x.op == MUL and is_equal(y, x.b, 1):
            return x.a
        if is_one(x):
            return unary(INV, y)
        if y.op == INV:
            return binary(MUL, x, y.a)
        if is_doubled(x) and is_doubled(y):
            return binary(DIV, x.a, y.a)
        if y.op == CONST and x.op == DIV and x.b.op == CONST and y.v * x.b.v == 1:
            return x.a
        if y.op == MUL and is_equal(y.b, x, 1):
            return unary(INV, y.a)
        if y.op == MUL and is_equal(y.a, x, 1):
            return unary(INV, y.b)
        if x.op == NEG:
            return unary(NEG, binary(DIV, x.a, y))
        if y.op == NEG:
            return unary(NEG, binary(DIV, x, y.a))
    elif op == POW:
        return power(x, y)
    elif op == FMIN or op == FMAX:
        if x is y:
            return x
    return N(op, x, y)


def power(x, y):
    """SXElem pow(): integer constant exponents are expanded by repeated squaring,
    0.5 becomes sqrt, other constant exponents become OP_CONSTPOW."""
    if x.op == CONST and y.op == CONST:
        return const(_safe_pow(x.v, y.v))
    if y.op == CONST:
        yv = y.v
        if yv == int(yv) and abs(yv) < 1e15:
            nn = int(yv)
            if nn == 0:
                return ONE
            if nn > 100 or nn < -100:
                return N(CONSTPOW, x, y)
            if nn < 0:
                return binary(DIV, ONE, power(x, const(-nn)))
            if nn % 2 == 1:
                return binary(MUL, x, power(x, const(nn - 1)))
            rt = power(x, const(nn // 2))
            return binary(MUL, rt, rt)
        if yv == 0.5:
            return unary(SQRT, x)
        return N(CONSTPOW, x, y)
    return N(POW, x, y)


# ---------------------------------------------------------------- derivative table (calculus.hpp)
def partials(n):
    """Return (d f/d a, d f/d b) of node ``n`` as nodes (b-part None for unary)."""
    op, a, b = n.op, n.a, n.b
    if op == ADD:
        return ONE, ONE
    if op == SUB:
        return ONE, MINUS_ONE
    if op == MUL:
        return b, a
    if op == DIV:
        # d/da = 1/b ; d/db = -f/b
        return binary(DIV, ONE, b), unary(NEG, binary(DIV, n, b))
    if op == CONSTPOW:
        # d/da = b*a^(b-1)
        return binary(MUL, b, power(a, const(b.v - 1.0))), ZERO
    if op == POW:
        return (binary(MUL, b, power(a, binary(SUB, b, ONE))),
                binary(MUL, unary(LOG, a), n))
    if op == ATAN2:
        t = binary(ADD, unary(SQ, a), unary(SQ, b))
        return binary(DIV, b, t), unary(NEG, binary(DIV, a, t))
    if op == FMIN:
        return N(LE, a, b), N(LT, b, a)   # casadi: d[0]=x<=y, d[1]=!d[0]
    if op == FMAX:
        return N(LE, b, a), N(LT, a, b)
    if op in (LT, LE, EQ, NE, AND, OR):
        return ZERO, ZERO
    if op == IF_ELSE_ZERO:
        return ZERO, N(NE, a, ZERO)
    if op == NEG:
        return MINUS_ONE, None
    if op == SQ:
        return binary(MUL, TWO, a), None
    if op == SQRT:
        return binary(DIV, ONE, binary(MUL, TWO, n)), None
    if op == SIN:
        return unary(COS, a), None
    if op == COS:
        return unary(NEG, unary(SIN, a)), None
    if op == TAN:
        return binary(DIV, ONE, unary(SQ, unary(COS, a))), None
    if op == ASIN:
        return binary(DIV, ONE, unary(SQRT, binary(SUB, ONE, unary(SQ, a)))), None
    if op == ACOS:
        return binary(DIV, MINUS_ONE, unary(SQRT, binary(SUB, ONE, unary(SQ, a)))), None
    if op == ATAN:
        return binary(DIV, ONE, binary(ADD, ONE, unary(SQ, a))), None
    if op == EXP:
        return n, None
    if op == LOG:
        return binary(DIV, ONE, a), None
    if op == FABS:
        return unary(SIGN, a), None
    if op == INV:
        return unary(NEG, unary(SQ, n)), None
    if op == TANH:
        return binary(SUB, ONE, unary(SQ, n)), None
    if op == SINH:
        return unary(COSH, a), None
    if op == COSH:
        return unary(SINH, a), None
    if op in (SIGN, NOT, FLOOR, CEIL):
        return ZERO, None
    raise NotImplementedError('partials for op %s' % op)


# ---------------------------------------------------------------- graph utilities
def topo(outputs):
    """Topologically sorted list of the non-leaf-and-leaf nodes reachable from ``outputs``
    (dependencies first), iterative to survive deep graphs."""
    order = []
    seen = set()
    for root in outputs:
        if id(root) in seen:
            continue
        stack = [(root, 0)]
        while stack:
            n, st = stack.pop()
            if st == 0:
                if id(n) in seen:
                    continue
                seen.add(id(n))
                stack.append((n, 1))
                if n.b is not None and id(n.b) not in seen:
                    stack.append((n.b, 0))
                if n.a is not None and id(n.a) not in seen:
                    stack.append((n.a, 0))
            else:
                order.append(n)
    return order


def to_str(n, depth=4):
    if n.op == CONST:
        return repr(n.v) if n.v != int(n.v) or abs(n.v) > 1e15 else str(int(n.v))
    if n.op == SYM:
        return n.name
    if depth <= 0:
        return '@%d' % n.id
    if n.op in BINARY:
        if n.op in (ADD, SUB, MUL, DIV):
            return '(%s%s%s)' % (to_str(n.a, depth - 1), OPNAME[n.op], to_str(n.b, depth - 1))
        return '%s(%s,%s)' % (OPNAME[n.op], to_str(n.a, depth - 1), to_str(n.b, depth - 1))
    return '%s(%s)' % (OPNAME[n.op], to_str(n.a, depth - 1))


def forward_jac(exprs, xs):
    """Symbolic Jacobian d exprs / d xs by sparse forward-mode AD over the expression graph.
    Returns a list (one per expr) of dicts {column index: derivative node}.
    (casadi picks forward or reverse mode per graph colouring; the result is the same function.)"""
    col = {id(x): j for j, x in enumerate(xs)}
    order = topo(exprs)
    D = {}
    for n in order:
        if n.op == SYM:
            j = col.get(id(n))
            if j is not None:
                D[id(n)] = {j: ONE}
            continue
        if n.op == CONST:
            continue
        da = D.get(id(n.a))
        db = D.get(id(n.b)) if n.b is not None else None
        if not da and not db:
            continue
        pa, pb = partials(n)
        d = {}
        if da:
            if is_one(pa):
                d.update(da)
            elif not is_zero(pa):
                for j, v in da.items():
                    d[j] = binary(MUL, pa, v)
        if db and pb is not None and not is_zero(pb):
            for j, v in db.items():
                t = v if is_one(pb) else binary(MUL, pb, v)
                if j in d:
                    d[j] = binary(ADD, d[j], t)
                else:
                    d[j] = t
        d = {j: v for j, v in d.items() if not is_zero(v)}
        if d:
            D[id(n)] = d
    return [D.get(id(e), {}) for e in exprs]


def reverse_gradient(root, xs):
    """Symbolic gradient d root / d xs by reverse-mode AD over the expression graph (what ``SX::gradient`` =
    ``jtimes(ex, arg, 1, true)`` does in casadi: ``SXFunction::ad_reverse`` walks the instruction list backwards and
    adds ``partial * seed`` to the adjoint of each dependency, with the construction-time simplifications applied).
    Returns a list (one per x) of nodes (ZERO where the root does not reach x)."""
    order = topo([root])
    bar = {id(root): ONE}
    for n in reversed(order):
        bn = bar.get(id(n))
        if bn is None or n.op == SYM or n.op == CONST or is_zero(bn):
            continue
        pa, pb = partials(n)
        for dep, pd in ((n.a, pa), (n.b, pb)):
            if dep is None or pd is None or dep.op == CONST or is_zero(pd):
                continue
            t = bn if is_one(pd) else binary(MUL, pd, bn)
            if is_zero(t):
                continue
            old = bar.get(id(dep))
            bar[id(dep)] = t if old is None else binary(ADD, old, t)
    return [bar.get(id(x), ZERO) for x in xs]
