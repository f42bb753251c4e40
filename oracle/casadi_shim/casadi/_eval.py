"""Numeric evaluation of SX graphs (TEST INFRASTRUCTURE — oracle only).

Restates what casadi's SXFunction virtual machine does: a topologically sorted instruction tape,
evaluated in FP64; structural Jacobian sparsity by dependency propagation (``sp_forward``); and the
compressed forward-mode Jacobian (one sweep per colour of a greedy distance-2 column colouring),
which is how casadi 3.6 forms ``jac_g`` for ``nlpsol``.  Vectorised over an arbitrary trailing
batch shape with numpy so that whole batches / all seed directions go through one pass.
"""
import numpy as np
from . import _sx

_UN = {
    _sx.NEG: np.negative, _sx.SQ: np.square, _sx.SQRT: np.sqrt, _sx.SIN: np.sin, _sx.COS: np.cos,
    _sx.TAN: np.tan, _sx.ASIN: np.arcsin, _sx.ACOS: np.arccos, _sx.ATAN: np.arctan, _sx.EXP: np.exp,
    _sx.LOG: np.log, _sx.FABS: np.abs, _sx.SIGN: np.sign, _sx.INV: np.reciprocal,
    _sx.NOT: lambda x: (x == 0).astype(float), _sx.FLOOR: np.floor, _sx.CEIL: np.ceil,
    _sx.TANH: np.tanh, _sx.SINH: np.sinh, _sx.COSH: np.cosh,
}
_BIN = {
    _sx.ADD: np.add, _sx.SUB: np.subtract, _sx.MUL: np.multiply, _sx.DIV: np.divide,
    _sx.POW: np.power, _sx.CONSTPOW: np.power, _sx.ATAN2: np.arctan2, _sx.FMIN: np.minimum,
    _sx.FMAX: np.maximum,
    _sx.LT: lambda x, y: (x < y).astype(float), _sx.LE: lambda x, y: (x <= y).astype(float),
    _sx.EQ: lambda x, y: (x == y).astype(float), _sx.NE: lambda x, y: (x != y).astype(float),
    _sx.AND: lambda x, y: ((x != 0) & (y != 0)).astype(float),
    _sx.OR: lambda x, y: ((x != 0) | (y != 0)).astype(float),
    _sx.IF_ELSE_ZERO: lambda x, y: np.where(x != 0, y, 0.0),
}


class Tape(object):
    def __init__(self, inputs, outputs):
        self.n_in = len(inputs)
        self.n_out = len(outputs)
        order = _sx.topo(outputs)
        slot = {}
        for j, s in enumerate(inputs):
            slot[id(s)] = j
        k = self.n_in
        consts = []
        instr = []
        for n in order:
            if id(n) in slot:
                continue
            if n.op == _sx.SYM:
                raise ValueError('free symbolic variable %s in Function' % n.name)
            slot[id(n)] = k
            if n.op == _sx.CONST:
                consts.append((k, n.v))
            elif n.b is None:
                instr.append((n.op, k, slot[id(n.a)], -1))
            else:
                instr.append((n.op, k, slot[id(n.a)], slot[id(n.b)]))
            k += 1
        self.n_work = k
        self.consts = consts
        self.instr = instr
        self.out_slots = [slot[id(o)] for o in outputs]
        self._inputs = inputs
        self._outputs = outputs

    # ---- plain evaluation; x: [n_in] or [n_in, B]
    def eval(self, x):
        x = np.asarray(x, dtype=float)
        w = [None] * self.n_work
        for j in range(self.n_in):
            w[j] = x[j]
        for k, v in self.consts:
            w[k] = v
        with np.errstate(all='ignore'):
            for op, k, a, b in self.instr:
                if b < 0:
                    w[k] = _UN[op](w[a])
                else:
                    w[k] = _BIN[op](w[a], w[b])
        out = [w[s] for s in self.out_slots]
        if x.ndim == 1:
            return np.array([float(v) for v in out])
        return np.stack([np.broadcast_to(np.asarray(v, dtype=float), x.shape[1:]) for v in out])

    # ---- structural dependency of every output on every input (bitsets as python ints)
    def pattern(self):
        dep = [0] * self.n_work
        for j in range(self.n_in):
            dep[j] = 1 << j
        for op, k, a, b in self.instr:
            dep[k] = dep[a] if b < 0 else (dep[a] | dep[b])
        rows = []
        for s in self.out_slots:
            bits = dep[s]
            cols = []
            j = 0
            while bits:
                low = bits & -bits
                cols.append(low.bit_length() - 1)
                bits ^= low
            rows.append(cols)
        return rows

    def jac_ccs(self, n_cols=None):
        """CCS (colind,row) of d out / d in[:n_cols] from the structural pattern."""
        n_cols = self.n_in if n_cols is None else n_cols
        rows = self.pattern()
        cols = [[] for _ in range(n_cols)]
        for i, cs in enumerate(rows):
            for j in cs:
                if j < n_cols:
                    cols[j].append(i)
        colind = [0]
        row = []
        for j in range(n_cols):
            row.extend(cols[j])
            colind.append(len(row))
        return np.array(colind, dtype=np.int64), np.array(row, dtype=np.int64)

    def colouring(self, colind, row):
        """greedy unidirectional (distance-2) column colouring."""
        n = len(colind) - 1
        row_cols = {}
        for j in range(n):
            for r in row[colind[j]:colind[j + 1]]:
                row_cols.setdefault(int(r), []).append(j)
        colour = [-1] * n
        for j in range(n):
            forbidden = set()
            for r in row[colind[j]:colind[j + 1]]:
                for j2 in row_cols[int(r)]:
                    if colour[j2] >= 0:
                        forbidden.add(colour[j2])
            c = 0
            while c in forbidden:
                c += 1
            colour[j] = c
        return colour

    def jac_values(self, x, colind=None, row=None):
        """Nonzeros of the Jacobian at x in CCS order, by compressed forward-mode sweeps."""
        if colind is None:
            colind, row = self.jac_ccs()
        colour = self.colouring(colind, row)
        nc = max(colour) + 1 if colour else 0
        x = np.asarray(x, dtype=float)
        # forward mode: value w, tangent dw [nc]
        w = [None] * self.n_work
        dw = [None] * self.n_work
        zero = np.zeros(nc)
        n_cols = len(colind) - 1
        for j in range(self.n_in):
            w[j] = x[j]
            if j < n_cols:
                s = np.zeros(nc)
                s[colour[j]] = 1.0
                dw[j] = s
            else:
                dw[j] = None
        for k, v in self.consts:
            w[k] = v
            dw[k] = None
        with np.errstate(all='ignore'):
            for op, k, a, b in self.instr:
                wa = w[a]
                da = dw[a]
                if b < 0:
                    f = _UN[op](wa)
                    w[k] = f
                    if da is None:
                        dw[k] = None
                    else:
                        dw[k] = _dun(op, wa, f) * da
                else:
                    wb = w[b]
                    db = dw[b]
                    f = _BIN[op](wa, wb)
                    w[k] = f
                    if da is None and db is None:
                        dw[k] = None
                    else:
                        pa, pb = _dbin(op, wa, wb, f)
                        if da is None:
                            dw[k] = pb * db
                        elif db is None:
                            dw[k] = pa * da
                        else:
                            dw[k] = pa * da + pb * db
        vals = np.zeros(len(row))
        outs = [dw[s] for s in self.out_slots]
        for j in range(n_cols):
            c = colour[j]
            for p in range(colind[j], colind[j + 1]):
                d = outs[int(row[p])]
                vals[p] = 0.0 if d is None else d[c]
        return vals


def _dun(op, a, f):
    if op == _sx.NEG: return -1.0
    if op == _sx.SQ: return 2.0 * a
    if op == _sx.SQRT: return 1.0 / (2.0 * f)
    if op == _sx.SIN: return np.cos(a)
    if op == _sx.COS: return -np.sin(a)
    if op == _sx.TAN: return 1.0 / np.square(np.cos(a))
    if op == _sx.ASIN: return 1.0 / np.sqrt(1.0 - a * a)
    if op == _sx.ACOS: return -1.0 / np.sqrt(1.0 - a * a)
    if op == _sx.ATAN: return 1.0 / (1.0 + a * a)
    if op == _sx.EXP: return f
    if op == _sx.LOG: return 1.0 / a
    if op == _sx.FABS: return np.sign(a)
    if op == _sx.INV: return -f * f
    if op == _sx.TANH: return 1.0 - f * f
    if op == _sx.SINH: return np.cosh(a)
    if op == _sx.COSH: return np.sinh(a)
    return 0.0


def _dbin(op, a, b, f):
    if op == _sx.ADD: return 1.0, 1.0
    if op == _sx.SUB: return 1.0, -1.0
    if op == _sx.MUL: return b, a
    if op == _sx.DIV: return 1.0 / b, -f / b
    if op == _sx.CONSTPOW: return b * np.power(a, b - 1.0), 0.0
    if op == _sx.POW: return b * np.power(a, b - 1.0), np.log(a) * f
    if op == _sx.ATAN2:
        t = a * a + b * b
        return b / t, -a / t
    if op == _sx.FMIN: return float(a <= b), float(not (a <= b))
    if op == _sx.FMAX: return float(a >= b), float(not (a >= b))
    if op == _sx.IF_ELSE_ZERO: return 0.0, float(a != 0)
    return 0.0, 0.0


def dependency_pattern(exprs, xs):
    t = Tape(xs, exprs)
    return t.pattern()
