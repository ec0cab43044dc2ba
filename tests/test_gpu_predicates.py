"""GPU parity, unit level: the device predicates against the oracle on the same seeded inputs (bit-exact)."""
import numpy as np
import pytest

from inmyroom_vulkan_b200 import scenes
from helpers import f32_bits

pytestmark = pytest.mark.gpu


def _rand_mats(rng, n, tscale=1.0, nonuniform=True):
    s = (rng.random((n, 3)) * 1.5 + 0.25) if nonuniform else np.ones((n, 3))
    return scenes.trs_matrices(rng.normal(size=(n, 3)) * tscale, scenes.random_quaternions(rng, n), s)


def test_sat_and_surface_bit_exact(gpu_ctx, oracle):
    rng = np.random.default_rng(11)
    n = 20000
    a = rng.normal(size=(n, 12)).astype(np.float32)
    b = rng.normal(size=(n, 12)).astype(np.float32)
    b[:, :3] *= 2.5
    # a slice of degenerate boxes: parallel axes (zero cross products => "overlap" on that axis, Paralgram.cpp:198-201)
    b[:500, 3:] = a[:500, 3:]
    a[500:600, 3:6] = 0
    mats = _rand_mats(rng, n)
    v, sa, sb = gpu_ctx.test_sat(a, b, mats)
    exp_v = np.array([oracle.sat(a[i], b[i], mats[i]) for i in range(n)], np.uint8)
    exp_sa = np.array([oracle.surface(a[i]) for i in range(n)], np.float32)
    exp_sb = np.array([oracle.surface(b[i], mats[i]) for i in range(n)], np.float32)
    assert np.array_equal(v, exp_v)
    assert 0.05 < exp_v.mean() < 0.95          # the sample exercises both verdicts
    assert np.array_equal(f32_bits(sa), f32_bits(exp_sa))
    assert np.array_equal(f32_bits(sb), f32_bits(exp_sb))


def test_tri_tri_bit_exact(gpu_ctx, oracle):
    rng = np.random.default_rng(12)
    n = 300000
    a = rng.normal(size=(n, 9)).astype(np.float32)
    b = (rng.normal(size=(n, 9)) * 0.8).astype(np.float32)
    m = _rand_mats(rng, 1, 0.2)[0]
    f, s = gpu_ctx.test_tri_tri(a, b, m)
    ef, es = oracle.tri_tri(a, b, m)
    assert np.array_equal(f, ef)
    assert (ef == 1).sum() > 1000
    assert np.array_equal(f32_bits(s), f32_bits(es))


def test_tri_tri_degenerate_cases(gpu_ctx, oracle):
    """coplanar, shared-edge, shared-vertex, zero-area and touching pairs."""
    rng = np.random.default_rng(13)
    A, B = [], []
    base = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float32)
    for k in range(2000):
        t = base * (0.5 + rng.random()) + rng.normal(size=3).astype(np.float32) * (k % 3 == 0)
        kind = k % 8
        if kind == 0:      # coplanar overlapping
            u = t + np.array([0.2, 0.1, 0], np.float32)
        elif kind == 1:    # coplanar disjoint
            u = t + np.array([5, 5, 0], np.float32)
        elif kind == 2:    # shared edge, different plane
            u = t.copy(); u[2] = t[2] + np.array([0, 0, 1], np.float32)
        elif kind == 3:    # shared vertex
            u = t.copy(); u[1] += np.array([0, 0.3, 1], np.float32); u[2] += np.array([0.4, 0, -1], np.float32)
        elif kind == 4:    # zero-area triangle
            u = np.stack([t[0], t[0], t[1]])
        elif kind == 5:    # touching: vertex on the other's plane within EPSILON
            u = t + np.array([0.1, 0.1, 5e-7], np.float32); u[2] += np.array([0, 0, 1], np.float32)
        elif kind == 6:    # identical
            u = t.copy()
        else:              # piercing
            u = np.array([[0.2, 0.2, -1], [0.3, 0.2, 1], [0.2, 0.3, 1]], np.float32) * (0.5 + rng.random())
        A.append(t.reshape(9)); B.append(u.reshape(9))
    A = np.array(A, np.float32); B = np.array(B, np.float32)
    f, s = gpu_ctx.test_tri_tri(A, B, None)
    ef, es = oracle.tri_tri(A, B, None)
    assert np.array_equal(f, ef)
    assert set(np.unique(ef).tolist()) >= {0, 1, 3}
    assert np.array_equal(f32_bits(s), f32_bits(es))


def test_pair_matrix_bit_exact(gpu_ctx, oracle):
    rng = np.random.default_rng(14)
    n = 4000
    a = _rand_mats(rng, n, 20.0); b = _rand_mats(rng, n, 20.0)
    # Sponza's node transform: rotation (0.7071,0,0,0.7071), non-uniform scale
    a[:100] = scenes.trs_matrices(np.zeros((100, 3)), np.tile([[0.70710678, 0, 0, 0.70710678]], (100, 1)),
                                  np.tile([[0.0399999991, 0.0400000028, 0.0400000028]], (100, 1)))
    out = gpu_ctx.test_pair_matrix(a, b)
    exp = np.stack([oracle.pair_matrix(a[i], b[i]) for i in range(n)])
    assert np.array_equal(f32_bits(out), f32_bits(exp))
