"""Restatement of ``casadi.tools`` structures (casadi 3.6.4, ``casadi/tools/structure3.py``) as far
as the reference uses them: ``entry``, ``struct``, ``struct_symSX``, ``struct_symMX``, ``struct_SX``,
``struct_MX`` and the numeric / symbolic struct views returned by calling a struct.
TEST INFRASTRUCTURE (oracle) — never imported by ``awebox_b200``.

Semantics restated here:
* entries are laid out in declaration order; every entry expands its ``repeat`` indices row-major
  and then its payload (a nested struct, or a column-major matrix of ``shape``);
* entries grouped in a tuple are interleaved at the level of their first repeat index;
* the canonical index of a scalar slot is ``(name, *repeat_indices, *sub_index)``, its label is
  ``'[' + ','.join(...) + ']'``;
* ``s[name, i, j, sub, ...]`` ("power index") walks that tree; missing / slice repeat indices give
  nested lists.
"""
from casadi import *          # noqa: F401,F403
from casadi import _as_mat, _Mat, _StructViewBase, _sx, _is_num, _shape_args
import casadi as _c
import numpy as _np


class entry(object):
    def __init__(self, name, **kw):
        self.name = name
        self.struct = kw.get('struct', None)
        rep = kw.get('repeat', [])
        if _is_num(rep):
            rep = [int(rep)]
        self.repeat = [int(r) for r in rep]
        self.expr = kw.get('expr', None)
        self.sym = kw.get('sym', None)
        shape = kw.get('shape', None)
        if self.struct is not None:
            if isinstance(self.struct, _View):
                self.struct = self.struct.struct
            self.shape = (self.struct.size, 1)
        elif shape is not None:
            if _is_num(shape):
                self.shape = (int(shape), 1)
            elif isinstance(shape, _c.Sparsity):
                self.shape = (shape.nrow, shape.ncol)
            else:
                self.shape = (int(shape[0]), int(shape[1])) if len(shape) == 2 else (int(shape[0]), 1)
        elif self.expr is not None and not isinstance(self.expr, list):
            e = _as_mat(self.expr)
            self.shape = e.shape
        elif self.expr is not None:
            first = self.expr
            while isinstance(first, list):
                first = first[0]
            self.shape = _as_mat(first).shape
            if not self.repeat:
                # infer the repeat structure from the nesting of the expression lists
                rep = []
                cur = self.expr
                while isinstance(cur, list):
                    rep.append(len(cur))
                    cur = cur[0]
                self.repeat = rep
        elif self.sym is not None:
            self.shape = _as_mat(self.sym).shape
        else:
            self.shape = (1, 1)
        self.elem_size = self.shape[0] * self.shape[1] if self.struct is None else self.struct.size
        self.offsets = {}

    def repeat_indices(self):
        import itertools
        return list(itertools.product(*[range(r) for r in self.repeat])) if self.repeat else [()]


class struct(object):
    def __init__(self, entries, order=None):
        self.entries = []
        groups = []
        import copy as _copy

        def _own(x):
            x = _copy.copy(x)
            x.offsets = {}
            return x
        for e in entries:
            if isinstance(e, (tuple, list)):
                grp = [_own(x) for x in e]
                self.entries.extend(grp)
                groups.append(grp)
            else:
                e = _own(e)
                self.entries.append(e)
                groups.append([e])
        self.dict = {}
        for e in self.entries:
            if e.name in self.dict:
                raise ValueError('duplicate struct entry %s' % e.name)
            self.dict[e.name] = e
        self.order = [tuple(x.name for x in g) if len(g) > 1 else g[0].name for g in groups]
        # layout
        self._canon = []
        off = 0
        for grp in groups:
            if len(grp) == 1 or all(not e.repeat for e in grp):
                for e in grp:
                    for mi in e.repeat_indices():
                        off = self._place(e, mi, off)
            else:
                kmax = max((e.repeat[0] if e.repeat else 1) for e in grp)
                for k in range(kmax):
                    for e in grp:
                        if e.repeat:
                            if k >= e.repeat[0]:
                                continue
                            import itertools
                            for rest in itertools.product(*[range(r) for r in e.repeat[1:]]):
                                off = self._place(e, (k,) + tuple(rest), off)
                        elif k == 0:
                            off = self._place(e, (), off)
        self.size = off
        self.shape = (off, 1)
        self._labels = None

    def _place(self, e, mi, off):
        e.offsets[mi] = off
        if e.struct is not None:
            for sub in e.struct._canon:
                self._canon.append((e.name,) + tuple(mi) + sub)
        else:
            for k in range(e.elem_size):
                self._canon.append((e.name,) + tuple(mi) + (k,))
        return off + e.elem_size

    # ---- queries
    def keys(self):
        return [e.name for e in self.entries]

    def __contains__(self, k):
        return k in self.dict

    def getCanonicalIndex(self, i):
        return self._canon[int(i)]

    def getLabel(self, i):
        return '[' + ','.join(str(k) for k in self._canon[int(i)]) + ']'

    def labels(self):
        if self._labels is None:
            self._labels = ['[' + ','.join(str(k) for k in c) + ']' for c in self._canon]
        return self._labels

    def getStruct(self, name):
        return self.dict[name].struct

    @property
    def struct(self):
        return self

    # ---- power index resolution
    def _resolve(self, keys, base=0):
        """-> leaf (flat index list, (n, m)) or nested python lists of leaves."""
        if not isinstance(keys, tuple):
            keys = (keys,)
        if len(keys) == 0:
            return (list(range(base, base + self.size)), (self.size, 1))
        name = keys[0]
        if name not in self.dict:
            raise KeyError('struct has no entry %r (keys %s)' % (name, self.keys()))
        e = self.dict[name]
        rest = list(keys[1:])
        sel = []
        for r in e.repeat:
            if rest and (isinstance(rest[0], (int, _np.integer, slice, list, range)) and not isinstance(rest[0], bool)):
                k = rest.pop(0)
                if isinstance(k, slice):
                    sel.append((list(range(*k.indices(r))), False))
                elif isinstance(k, (list, range)):
                    sel.append(([int(i) % r for i in k], False))
                else:
                    k = int(k)
                    if k < 0:
                        k += r
                    if k < 0 or k >= r:
                        raise IndexError('repeat index %d out of range %d for entry %s' % (k, r, name))
                    sel.append(([k], True))
            elif rest and isinstance(rest[0], _Mat):
                k = int(float(rest.pop(0)))
                sel.append(([k % r], True))
            else:
                sel.append((list(range(r)), False))
        rest = tuple(rest)

        def leaf(mi):
            off = base + e.offsets[tuple(mi)]
            if e.struct is not None:
                if rest:
                    return e.struct._resolve(rest, off)
                return (list(range(off, off + e.elem_size)), (e.elem_size, 1))
            n, m = e.shape
            idx = list(range(off, off + n * m))
            if rest:
                im = _c.DM._make(n, m, [float(i) for i in idx])
                sub = im[rest[0]] if len(rest) == 1 else im[rest[0], rest[1]]
                return ([int(v) for v in sub.e], sub.shape)
            return (idx, (n, m))

        def build(d, mi):
            if d == len(sel):
                return leaf(mi)
            idxs, scalar = sel[d]
            if scalar:
                return build(d + 1, mi + [idxs[0]])
            return [build(d + 1, mi + [i]) for i in idxs]

        return build(0, [])


def _map_leaves(tree, f):
    if isinstance(tree, list):
        return [_map_leaves(t, f) for t in tree]
    return f(tree)


def _flatten_leaves(tree, out):
    if isinstance(tree, list):
        for t in tree:
            _flatten_leaves(t, out)
    else:
        out.extend(tree[0])
    return out


class _FlatIndexer(object):
    def __init__(self, s):
        self._s = s

    def __getitem__(self, keys):
        return _flatten_leaves(self._s.struct._resolve(keys), [])


class _MatIndexer(object):
    def __init__(self, s):
        self._s = s

    def __getitem__(self, keys):
        return _map_leaves(self._s.struct._resolve(keys),
                           lambda lf: _c.DM._make(lf[1][0], lf[1][1], [float(i) for i in lf[0]]))


class _Prefix(object):
    def __init__(self, owner, pre=()):
        self._o = owner
        self._pre = tuple(pre)

    def __getitem__(self, keys):
        if not isinstance(keys, tuple):
            keys = (keys,)
        return _PrefixedView(self._o, self._pre + keys)


class _PrefixedView(object):
    def __init__(self, owner, pre):
        self._o = owner
        self._pre = pre

    def __getitem__(self, keys):
        if not isinstance(keys, tuple):
            keys = (keys,)
        return self._o[self._pre + keys]

    def __setitem__(self, keys, v):
        if not isinstance(keys, tuple):
            keys = (keys,)
        self._o[self._pre + keys] = v

    @property
    def prefix(self):
        return _Prefix(self._o, self._pre)

    def keys(self):
        s = self._o.struct
        for k in self._pre:
            if isinstance(k, str):
                s = s.dict[k].struct
        return s.keys()


class _View(_StructViewBase):
    """A struct together with a master vector (numeric DM or symbolic SX): DMStruct / SXStruct /
    MXStruct of casadi.tools."""

    def __init__(self, struct_, master):
        self._struct = struct_
        self.master = master

    @property
    def struct(self):
        return self._struct

    @property
    def cat(self):
        return self.master

    @property
    def size(self):
        return self._struct.size

    @property
    def shape(self):
        return self._struct.shape

    def keys(self):
        return self._struct.keys()

    def labels(self):
        return self._struct.labels()

    def getCanonicalIndex(self, i):
        return self._struct.getCanonicalIndex(i)

    def getLabel(self, i):
        return self._struct.getLabel(i)

    @property
    def f(self):
        return _FlatIndexer(self)

    @property
    def i(self):
        return _MatIndexer(self)

    @property
    def prefix(self):
        return _Prefix(self)

    @property
    def entries(self):
        return self._struct.entries

    @property
    def order(self):
        return self._struct.order

    @property
    def dict(self):
        return self._struct.dict

    def __contains__(self, k):
        return k in self._struct.dict

    def __getitem__(self, keys):
        m = self.master
        return _map_leaves(self._struct._resolve(keys),
                           lambda lf: m._make(lf[1][0], lf[1][1], [m.e[i] for i in lf[0]]))

    def __setitem__(self, keys, val):
        tree = self._struct._resolve(keys)
        self._assign(tree, val)

    def _assign(self, tree, val):
        if isinstance(tree, list):
            if isinstance(val, (list, tuple)) and len(val) == len(tree):
                for t, v in zip(tree, val):
                    self._assign(t, v)
            else:
                for t in tree:
                    self._assign(t, val)
            return
        idx, shape = tree
        v = self.master._coerce_like(val)
        if v.numel() == 1:
            for i in idx:
                self.master.e[i] = v.e[0]
        else:
            if v.numel() != len(idx):
                raise ValueError('struct assignment: size %d into %d slots' % (v.numel(), len(idx)))
            for i, x in zip(idx, v.e):
                self.master.e[i] = x

    def __call__(self, arg=0.0):
        return self._struct_call(arg)

    def _struct_call(self, arg):
        return _call_struct(self._struct, arg)

    def __deepcopy__(self, memo):
        import copy
        return _View(self._struct, copy.deepcopy(self.master))

    def __array__(self, dtype=None, copy=None):
        return _np.array(self.master, dtype=dtype)

    def full(self):
        return self.master.full()

    def __repr__(self):
        return 'StructView(%d)' % self._struct.size


def _call_struct(s, arg):
    if isinstance(arg, _View):
        arg = arg.master
    if isinstance(arg, _c.Function) or callable(arg) and not isinstance(arg, _Mat):
        raise TypeError('struct(...) called with a callable')
    if _is_num(arg):
        return _View(s, _c.DM._make(s.size, 1, [float(arg)] * s.size))
    m = _as_mat(arg)
    if m.numel() == 1 and s.size != 1:
        return _View(s, m._make(s.size, 1, [m.e[0]] * s.size))
    if m.numel() != s.size:
        raise ValueError('struct(...): vector of length %d for a struct of size %d' % (m.numel(), s.size))
    return _View(s, m._make(s.size, 1, list(m.e)))


class _SymStruct(_View):
    """struct_symSX / struct_symMX: a struct whose master vector is made of fresh symbols."""

    def __init__(self, entries, order=None):
        s = struct(entries)
        syms = [_sx.sym('_'.join(str(k) for k in c)) for c in s._canon]
        _View.__init__(self, s, _c.SX._make(s.size, 1, syms))

    def __deepcopy__(self, memo):
        return self


class struct_symSX(_SymStruct):
    pass


class struct_symMX(_SymStruct):
    pass


class struct_SX(_View):
    """A struct whose slots hold given expressions (``entry(name, expr=...)``)."""

    def __init__(self, entries, order=None):
        s = struct(entries)
        master = _c.SX.zeros(s.size, 1)
        for e in s.entries:
            for mi in e.repeat_indices():
                ex = e.expr
                for k in mi:
                    ex = ex[k]
                ex = _as_mat(ex)
                if not isinstance(ex, _c.SX):
                    ex = _c.SX(ex)
                if ex.numel() != e.elem_size:
                    raise ValueError('struct_SX entry %s: expression size mismatch' % e.name)
                off = e.offsets[tuple(mi)]
                for k, v in enumerate(ex.e):
                    master.e[off + k] = v
        _View.__init__(self, s, master)


struct_MX = struct_SX
DMStruct = _View
SXStruct = _View
MXStruct = _View


def struct_load(*a, **k):
    raise NotImplementedError
