"""Radau / Legendre collocation constants (awebox/ocp/collocation.py:67-144, 185-200).

``coeff_collocation[r, j] = l_r'(tau_j)`` and ``coeff_continuity[r] = l_r(1)`` for the Lagrange basis
on ``tau = [0, collocation_points(d, scheme)]``.  The reference obtains the points from CasADi's
long-double table and differentiates the basis symbolically; here the points are the roots of the
corresponding Jacobi polynomials computed in 50-digit arithmetic (mpmath) and the basis is
differentiated in closed form in the same arithmetic, then rounded once to FP64.
"""
import numpy as np

try:
    import mpmath
except ImportError:          # pragma: no cover - mpmath ships with sympy in this image
    mpmath = None


def collocation_points(d, scheme='radau'):
    mpmath.mp.dps = 50
    if scheme == 'radau':
        # right Radau IIA nodes on (0,1]: roots of P_{d-1}(s) - P_d(s), s = 2 tau - 1
        f = lambda t: mpmath.legendre(d - 1, 2 * t - 1) - mpmath.legendre(d, 2 * t - 1)
    elif scheme == 'legendre':
        f = lambda t: mpmath.legendre(d, 2 * t - 1)
    else:
        raise ValueError('unknown collocation scheme %r' % scheme)
    roots = []
    n_grid = 2000
    prev_t, prev_v = mpmath.mpf(0), f(mpmath.mpf(0))
    for i in range(1, n_grid + 1):
        t = mpmath.mpf(i) / n_grid
        v = f(t)
        if prev_v * v < 0:
            roots.append(mpmath.findroot(f, (prev_t, t), solver='anderson', tol=1e-45))
        prev_t, prev_v = t, v
    if scheme == 'radau':
        roots.append(mpmath.mpf(1))
    if len(roots) != d:
        raise RuntimeError('found %d collocation points for order %d' % (len(roots), d))
    return roots


class Collocation(object):
    def __init__(self, n_k, d, scheme='radau'):
        self.n_k, self.d, self.scheme = int(n_k), int(d), scheme
        mp = mpmath.mpf
        tau = [mp(0)] + collocation_points(self.d, scheme)
        self.tau_root = np.array([float(t) for t in tau])
        C = np.zeros((self.d + 1, self.d + 1))
        D = np.zeros(self.d + 1)
        for j in range(self.d + 1):
            others = [r for r in range(self.d + 1) if r != j]
            denom = mp(1)
            for r in others:
                denom *= (tau[j] - tau[r])
            # l_j(1)
            num = mp(1)
            for r in others:
                num *= (mp(1) - tau[r])
            D[j] = float(num / denom)
            # l_j'(tau_r)
            for r in range(self.d + 1):
                s = mp(0)
                for a in others:
                    prod = mp(1)
                    for b in others:
                        if b != a:
                            prod *= (tau[r] - tau[b])
                    s += prod
                C[j][r] = float(s / denom)
        self.coeff_collocation = C      # [r, j]
        self.coeff_continuity = D
        self.quad_weights = np.linalg.solve(C[1:, 1:], D[1:])
