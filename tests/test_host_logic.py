"""Host-side pieces: expression graph, collocation constants, layout, synthetic points, sharding."""
import numpy as np
import pytest

from conftest import get_ocp
from awebox_b200 import sym as S


def test_sym_jacobian_against_central_differences():
    x = S.symvec('x', 4)
    f = [S.sqrt(x[0] * x[0] + x[1] * x[1] + 1e-3) * S.power(x[2], 2.5) / (x[3] + 2.0),
         S.atan2(x[0], x[1]) * S.exp(x[2]) - S.log(x[3] + 3.0) * S.sin(x[0]),
         S.power(x[0] + 2.0, x[1]) + S.cos(x[2]) * S.tan(0.1 * x[3])]
    J = S.jacobian_dense(f, x)
    pt = np.array([0.3, 0.7, 1.1, 0.4])
    Jn = S.evaluate(list(J.reshape(-1)), x, pt).reshape(3, 4)
    h = 1e-6
    for j in range(4):
        e = np.zeros(4)
        e[j] = h
        fd = (S.evaluate(f, x, pt + e) - S.evaluate(f, x, pt - e)) / (2 * h)
        assert np.allclose(Jn[:, j], fd, rtol=1e-7, atol=1e-9)


def test_sym_structural_dependency_and_cse():
    x = S.symvec('y', 3)
    a = x[0] * x[1]
    assert a is x[1] * x[0]                      # hash-consed, commutative canonical form
    assert (a - a).is_zero() and (a / a).is_one()
    assert S.dependency([a + x[2] * 0.0, x[2]], x) == [[0, 1], [2]]


def test_collocation_constants():
    from awebox_b200.ocp.collocation import Collocation
    c = Collocation(10, 4, 'radau')
    assert np.allclose(c.tau_root, [0.0, 0.0885879595127039, 0.409466864440735, 0.787659461760847, 1.0], atol=1e-15)
    assert np.allclose(c.coeff_collocation.sum(axis=0), 0.0, atol=1e-12)   # derivative of a constant
    assert np.array_equal(c.coeff_continuity, [0.0, 0.0, 0.0, 0.0, 1.0])    # Radau: tau_d = 1
    assert abs(c.quad_weights.sum() - 1.0) < 1e-13
    t = c.tau_root
    for p in range(1, 5):                         # exact differentiation of polynomials up to degree d
        assert np.allclose(c.coeff_collocation.T @ t ** p, p * t ** (p - 1), atol=1e-12)


def test_layout_is_the_contiguous_interval_layout():
    ocp = get_ocp('ampyx_ap2', 4)
    for k in range(ocp.n_k):
        assert ocp.v_u[k] == ocp.v_x[k] + ocp.n_x
        assert ocp.v_xdot[k] == ocp.v_u[k] + ocp.n_u
        assert ocp.v_z[k] == ocp.v_xdot[k] + ocp.n_x
        assert ocp.v_cx[k][0] == ocp.v_z[k] + ocp.n_z
    assert ocp.p_theta0 == ocp.n_v + ocp.n_var + 20
    assert ocp.colind[-1] == ocp.nnz and np.all(np.diff(ocp.colind) >= 0)
    for c in range(ocp.n_v):                      # rows strictly increasing inside every column
        r = ocp.row[ocp.colind[c]:ocp.colind[c + 1]]
        assert np.all(np.diff(r) > 0)


def test_synthetic_points_are_finite_and_seeded():
    from awebox_b200 import synthetic
    ocp = get_ocp('ampyx_ap2', 4)
    W1 = synthetic.batch(ocp, 5, seed=3)
    W2 = synthetic.batch(ocp, 5, seed=3)
    assert W1.shape == (5, ocp.n_v) and np.array_equal(W1, W2) and np.all(np.isfinite(W1))
    w0 = synthetic.base_point(ocp)
    R = w0[ocp.v_x[0] + 9:ocp.v_x[0] + 18].reshape(3, 3).T
    assert np.allclose(R.T @ R, np.eye(3), atol=1e-12)


def test_shard_ranges_cover_the_batch():
    from awebox_b200.parallel import shard_range, shard_sizes
    for B in (0, 1, 7, 64, 4096, 8191):
        for world in (1, 2, 3, 8):
            ranges = [shard_range(B, r, world) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == B
            assert all(ranges[i][1] == ranges[i + 1][0] for i in range(world - 1))
            assert max(shard_sizes(B, world)) - min(shard_sizes(B, world)) <= 1
