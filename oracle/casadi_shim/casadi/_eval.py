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


# ------------------------------------------------------------------ Hessian of the Lagrangian (nlpsol's hess_l)
class LagrangianHessian(object):
    """``triu(hessian(lam_g' g(x, p), x))`` the way casadi's ``nlpsol`` builds ``nlp_hess_l`` for an SX oracle:
    the gradient of the scalar by symbolic reverse mode (``_sx.reverse_gradient``), its Jacobian pattern by structural
    dependency propagation over the gradient graph (symmetric by construction), the values by forward tangents of
    the gradient.  Compression of the forward sweeps: columns that couple to (almost) everything (``theta``,
    ``phi``) get a direction each, the rest a greedy distance-2 colouring of the pattern without those rows and
    columns; entries in the dense rows are read from the symmetric position."""

    def __init__(self, g_inputs, g_outputs, n_x, f_output=None):
        self.n_x = n_x
        self.n_in = len(g_inputs)
        self.n_g = len(g_outputs)
        self.lam = [_sx.sym('lam_g_%d' % i) for i in range(self.n_g)]
        L = _sx.ZERO
        for l, g in zip(self.lam, g_outputs):
            L = _sx.binary(_sx.ADD, L, _sx.binary(_sx.MUL, l, g))
        self.lam_f = []
        if f_output is not None:
            # gamma = lam_f f + lam_g' g (nlpsol: "triu:hess:gamma:x:x" with gamma = {f, g})
            self.lam_f = [_sx.sym('lam_f')]
            L = _sx.binary(_sx.ADD, _sx.binary(_sx.MUL, self.lam_f[0], f_output), L)
        grad = _sx.reverse_gradient(L, list(g_inputs[:n_x]))
        self.tape = Tape(list(g_inputs) + self.lam + self.lam_f, grad)
        rows = self.tape.pattern()                  # row i: columns (inputs) d grad_i depends on
        cols = [[] for _ in range(n_x)]
        for i, cs in enumerate(rows):
            for j in cs:
                if j < n_x:
                    cols[j].append(i)
        # symmetric closure (the structural pattern of a reverse-over-forward graph is symmetric up to
        # simplifications; casadi symmetrises), then the upper triangle
        full = [set(c) for c in cols]
        for j in range(n_x):
            for i in list(full[j]):
                full[i].add(j)
        self.full_cols = [sorted(c) for c in full]
        colind, row = [0], []
        for j in range(n_x):
            row.extend(i for i in self.full_cols[j] if i <= j)
            colind.append(len(row))
        self.colind, self.row = np.array(colind, dtype=np.int64), np.array(row, dtype=np.int64)
        self._colour()

    def _colour(self):
        n = self.n_x
        deg = np.array([len(c) for c in self.full_cols])
        limit = max(64, 4 * int(np.median(deg[deg > 0])) if (deg > 0).any() else 64)
        self.head = [j for j in range(n) if deg[j] > limit]
        is_head = np.zeros(n, dtype=bool)
        is_head[self.head] = True
        colour = [-1] * n
        for c, j in enumerate(self.head):
            colour[j] = c
        base = len(self.head)
        row_cols = {}
        for j in range(n):
            if is_head[j]:
                continue
            for r in self.full_cols[j]:
                if not is_head[r]:
                    row_cols.setdefault(r, []).append(j)
        for j in range(n):
            if is_head[j]:
                continue
            forbidden = set()
            for r in self.full_cols[j]:
                if is_head[r]:
                    continue
                for j2 in row_cols[r]:
                    if colour[j2] >= 0:
                        forbidden.add(colour[j2])
            c = base
            while c in forbidden:
                c += 1
            colour[j] = c
        self.colour, self.is_head = colour, is_head
        self.n_colours = max(colour) + 1 if colour else 0

    def tangent(self, xin, seeds):
        """forward tangents of the gradient tape: d grad / d inputs . seeds;  seeds {input slot: array [nd]}"""
        t = self.tape
        w = [None] * t.n_work
        dw = [None] * t.n_work
        for j in range(t.n_in):
            w[j] = xin[j]
            dw[j] = seeds.get(j)
        for k, v in t.consts:
            w[k] = v
        with np.errstate(all='ignore'):
            for op, k, a, b in t.instr:
                wa, da = w[a], dw[a]
                if b < 0:
                    f = _UN[op](wa)
                    w[k] = f
                    dw[k] = None if da is None else _dun(op, wa, f) * da
                else:
                    wb, db = w[b], dw[b]
                    f = _BIN[op](wa, wb)
                    w[k] = f
                    if da is None and db is None:
                        continue
                    pa, pb = _dbin(op, wa, wb, f)
                    if da is None:
                        dw[k] = pb * db
                    elif db is None:
                        dw[k] = pa * da
                    else:
                        dw[k] = pa * da + pb * db
        return [w[s] for s in t.out_slots], [dw[s] for s in t.out_slots]

    def values(self, x, p, lam, lam_f=None):
        """non-zeros of triu(H) in CCS order at (x, p, lam)"""
        xin = np.concatenate([x, p, lam] + ([[lam_f]] if self.lam_f else []))
        nc = self.n_colours
        seeds = {}
        for j in range(self.n_x):
            s = np.zeros(nc)
            s[self.colour[j]] = 1.0
            seeds[j] = s
        _, dg = self.tangent(xin, seeds)
        vals = np.zeros(len(self.row))
        for j in range(self.n_x):
            for q in range(self.colind[j], self.colind[j + 1]):
                i = int(self.row[q])
                # entry (i, j), i <= j: column j seeded alone within row i unless row i is dense
                if self.is_head[i] and not self.is_head[j]:
                    d = dg[j]
                    vals[q] = 0.0 if d is None else d[self.colour[i]]
                else:
                    d = dg[i]
                    vals[q] = 0.0 if d is None else d[self.colour[j]]
        return vals

    def times(self, x, p, lam, V, lam_f=None):
        """H . V for a dense direction block V [n_x, nd] (full symmetric product; independent of the colouring)"""
        xin = np.concatenate([x, p, lam] + ([[lam_f]] if self.lam_f else []))
        seeds = {j: V[j] for j in range(self.n_x)}
        _, dg = self.tangent(xin, seeds)
        out = np.zeros((self.n_x, V.shape[1]))
        for i, d in enumerate(dg):
            if d is not None:
                out[i] = d
        return out
