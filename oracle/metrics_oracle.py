"""TEST INFRASTRUCTURE ONLY - numpy restatement of COV / MMD / KNN / f_score / JSD
(lib/networks/utils.py:38-144) used to check the host-side metric helpers."""
import numpy as np


def cov(d, axis=1):
    return len(np.unique(d.argmin(axis))) / float(d.shape[axis])


def mmd(d, axis=1):
    return float(d.min((axis + 1) % 2).mean())


def knn1(Mxx, Mxy, Myy):
    n0, n1 = Mxx.shape[0], Myy.shape[0]
    label = np.concatenate([-np.ones(n0), np.ones(n1)])
    M = np.block([[Mxx, Mxy], [Mxy.T, Myy]]).astype(np.float64)
    np.fill_diagonal(M, np.inf)
    pred = label[M.argmin(0)]
    pred = np.where(pred >= 0, 1.0, -1.0)
    return float((label == pred).mean())


def f_score(ld, rd, threshold=0.001):
    precision = 100. * (rd < threshold).mean(1)
    recall = 100. * (ld < threshold).mean(1)
    return 2. * precision * recall / (precision + recall + 1e-7)


def jsd(c1, c2, res=28):
    def occ(c):
        p = c.reshape(-1, 3)
        i = np.floor((p + 0.5) * res).astype(np.int64)
        ok = np.all((p >= -0.5) & (p < 0.5), axis=1) & np.all((i >= 0) & (i < res), axis=1)
        h = np.zeros((res, res, res))
        np.add.at(h, (i[ok, 0], i[ok, 1], i[ok, 2]), 1.0)
        return (h / h.sum()).ravel()

    def ent(p):
        p = p[p > 0]
        return -(p * np.log2(p)).sum()
    a, b = occ(c1), occ(c2)
    return ent((a + b) / 2) - 0.5 * (ent(a) + ent(b))
