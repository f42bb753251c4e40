"""Make the reference importable in THIS container (TEST INFRASTRUCTURE — oracle only).

/root/reference is pure Python whose arithmetic is executed by casadi==3.6.4, which is absent here.
``activate()`` puts the CasADi restatement (oracle/casadi_shim) and /root/reference on sys.path and
installs inert stand-ins for the presentation-only dependencies the reference imports at module
import time (matplotlib & co; awebox/__init__.py imports awebox.viz).  Nothing here is used on the
GPU box or by the product package.
"""
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE = os.environ.get('AWEBOX_REFERENCE', '/root/reference')


class _Inert(types.ModuleType):
    """module stand-in: every attribute is an inert callable/class."""
    __path__ = []

    def __getattr__(self, name):
        if name.startswith('__'):
            raise AttributeError(name)
        full = self.__name__ + '.' + name
        if full in sys.modules:
            return sys.modules[full]
        return _inert_obj


class _InertObj(object):
    def __call__(self, *a, **k):
        return self

    def __getattr__(self, name):
        if name.startswith('__'):
            raise AttributeError(name)
        return self

    def __iter__(self):
        return iter(())

    def __mro_entries__(self, bases):
        return (object,)


_inert_obj = _InertObj()

_STUB_ROOTS = ('matplotlib', 'mpl_toolkits', 'reportlab')


class _InertFinder(object):
    """meta-path finder: any (sub)module of the presentation-only packages is an inert module."""

    def find_spec(self, fullname, path=None, target=None):
        import importlib.machinery
        if fullname.split('.')[0] in _STUB_ROOTS:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        return _Inert(spec.name)

    def exec_module(self, module):
        pass


def available():
    return os.path.isdir(os.path.join(REFERENCE, 'awebox'))


def real_casadi_requested():
    """AWEBOX_REAL_CASADI=1: run the reference over the REAL casadi package (a machine that has casadi==3.6.4 and
    the reference checked out; not this container) instead of the restatement under oracle/casadi_shim."""
    return os.environ.get('AWEBOX_REAL_CASADI', '') not in ('', '0')


def activate():
    if not available():
        raise RuntimeError('reference tree not present at %s' % REFERENCE)
    if real_casadi_requested():
        if REFERENCE not in sys.path:
            sys.path.insert(0, REFERENCE)
        if not any(isinstance(f, _InertFinder) for f in sys.meta_path):
            sys.meta_path.append(_InertFinder())
        import casadi
        if 'shim' in casadi.__version__:
            raise RuntimeError('AWEBOX_REAL_CASADI=1 but the imported casadi is the restatement')
        return casadi
    shim = os.path.join(HERE, 'casadi_shim')
    if shim not in sys.path:
        sys.path.insert(0, shim)
    if REFERENCE not in sys.path:
        sys.path.insert(1, REFERENCE)
    if not any(isinstance(f, _InertFinder) for f in sys.meta_path):
        sys.meta_path.append(_InertFinder())
    import casadi
    assert 'shim' in casadi.__version__
    return casadi
