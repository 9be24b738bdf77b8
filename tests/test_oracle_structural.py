"""CPU: the C restatement of the reference's NN-distance kernel agrees with an independent
numpy brute force (values to fp32 rounding, indices exactly away from ties), incl. ragged sizes,
tie-breaking and the chunk-of-512 structure."""
import numpy as np
import pytest

from oracle import structural as so


@pytest.mark.parametrize("b,n,m", [(2, 64, 64), (1, 513, 1030), (3, 7, 1), (1, 1, 5), (2, 1200, 37)])
def test_nndistance_vs_bruteforce(b, n, m):
    rng = np.random.default_rng(n * 1000 + m)
    x = rng.uniform(-0.5, 0.5, (b, n, 3)).astype(np.float32)
    y = rng.uniform(-0.5, 0.5, (b, m, 3)).astype(np.float32)
    d1, i1, d2, i2 = so.nndistance(x, y)
    bd1, bi1, bd2, bi2 = so.brute_nn(x, y)
    np.testing.assert_allclose(d1, bd1, rtol=1e-6, atol=1e-12)
    np.testing.assert_allclose(d2, bd2, rtol=1e-6, atol=1e-12)
    # indices: identical except where fp32 rounding creates/breaks a tie
    assert (i1 == bi1).mean() > 0.999 and (i2 == bi2).mean() > 0.999


def test_ties_lowest_index_wins():
    x = np.zeros((1, 4, 3), np.float32)
    y = np.zeros((1, 1100, 3), np.float32)
    y[0, :, 0] = 1.0
    y[0, 700, 0] = 0.5  # unique best in the second 512-chunk
    d1, i1, d2, i2 = so.nndistance(x, y)
    assert (i1 == 700).all() and np.allclose(d1, 0.25)
    y[0, 700, 0] = 1.0  # all equal -> index 0
    d1, i1, _, _ = so.nndistance(x, y)
    assert (i1 == 0).all()
    assert (i2 == 0).all()


def test_pairwise_cd_matches_definition():
    rng = np.random.default_rng(0)
    A = rng.uniform(-0.5, 0.5, (3, 50, 3)).astype(np.float32)
    B = rng.uniform(-0.5, 0.5, (4, 60, 3)).astype(np.float32)
    out = so.pairwise_cd(A, B)
    for i in range(3):
        for j in range(4):
            bd1, _, bd2, _ = so.brute_nn(A[i:i + 1], B[j:j + 1])
            assert abs(out[i, j] - (bd1.mean() + bd2.mean())) < 1e-6
    G = so.pairwise_cd(A, A)
    assert np.array_equal(G, G.T) and np.all(np.diag(G) == 0)


def test_nndistance_grad_matches_finite_difference():
    rng = np.random.default_rng(1)
    x = rng.uniform(-0.5, 0.5, (1, 20, 3)).astype(np.float32)
    y = rng.uniform(-0.5, 0.5, (1, 30, 3)).astype(np.float32)
    d1, i1, d2, i2 = so.nndistance(x, y)
    g1, g2 = so.nndistance_grad(x, y, i1, i2, np.ones_like(d1), np.ones_like(d2))

    def f(xx, yy):
        a, _, b, _ = so.brute_nn(xx, yy)
        return a.sum() + b.sum()
    eps = 1e-3
    xp = x.copy(); xp[0, 3, 1] += eps
    xm = x.copy(); xm[0, 3, 1] -= eps
    fd = (f(xp, y) - f(xm, y)) / (2 * eps)
    assert abs(fd - g1[0, 3, 1]) < 1e-2 * max(1, abs(fd))
