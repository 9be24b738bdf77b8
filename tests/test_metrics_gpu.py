"""GPU: the fused generation-score reductions (dpf_cd_scores) and the voxel histogram behind JSD
(dpf_voxel_hist) against the numpy oracle and the reference-API torch functions (lib/networks/utils.py:45-144)."""
import numpy as np
import pytest
import torch

from oracle import metrics_oracle as mo

pytestmark = pytest.mark.gpu


def _mats(S1, S2, seed, dev):
    rng = np.random.default_rng(seed)
    gg = np.abs(rng.normal(size=(S1, S1))).astype(np.float32); gg = (gg + gg.T) / 2; np.fill_diagonal(gg, 0)
    tt = np.abs(rng.normal(size=(S2, S2))).astype(np.float32); tt = (tt + tt.T) / 2; np.fill_diagonal(tt, 0)
    gt = np.abs(rng.normal(size=(S1, S2))).astype(np.float32)
    return gg, gt, tt


@pytest.mark.parametrize("S1,S2", [(1, 1), (2, 3), (40, 40), (257, 130), (1000, 1000)])
def test_cd_scores_vs_oracle_and_reference_api(cuda, native_lib, S1, S2):
    from dpf_nets_b200.lib.networks import utils as u
    from dpf_nets_b200.ops.metrics import cd_scores
    gg, gt, tt = _mats(S1, S2, S1 * 7 + S2, cuda)
    G, X, T = (torch.from_numpy(a).to(cuda) for a in (gg, gt, tt))
    cov, mmd, nna = cd_scores(G, X, T).tolist()
    assert cov == pytest.approx(mo.cov(gt), abs=1e-7)
    assert mmd == pytest.approx(mo.mmd(gt), rel=1e-6)
    assert nna == pytest.approx(mo.knn1(gg, gt, tt), abs=1e-7)
    # the reference-API functions (torch ops with the reference's return conventions) agree
    assert cov == pytest.approx(u.COV(X), abs=1e-7)
    assert mmd == pytest.approx(u.MMD(X), rel=1e-6)
    if S1 + S2 > 2:
        assert nna == pytest.approx(u.KNN(G, X, T, 1), abs=1e-7)


def test_cd_scores_ties_lowest_index(cuda, native_lib):
    from dpf_nets_b200.ops.metrics import cd_scores
    # all distances equal: every item's nearest neighbour is index 0 of the concatenated order (a generated one,
    # except for item 0 itself whose nearest other item is generated item 1) -> gen items correct, ref items wrong
    S = 5
    one = torch.ones((S, S), device=cuda)
    cov, mmd, nna = cd_scores(one, one, one).tolist()
    assert cov == pytest.approx(1.0 / S) and mmd == pytest.approx(1.0) and nna == pytest.approx(0.5)


@pytest.mark.parametrize("shape", [(6, 300, 3), (2, 5, 3), (50, 2048, 3)])
def test_voxel_hist_and_jsd_vs_reference_binning(cuda, native_lib, shape):
    from dpf_nets_b200.lib.networks import utils as u
    from dpf_nets_b200.ops.metrics import jsd_from_hists, voxel_hist
    rng = np.random.default_rng(shape[0])
    c1 = rng.uniform(-0.6, 0.6, shape).astype(np.float32)                  # some points outside the cube
    c2 = (rng.normal(size=shape) * 0.2).astype(np.float32)
    edges = (-0.5 + np.arange(29) * (1. / 28))
    c1.reshape(-1)[:29 if c1.size >= 29 else 0] = edges[:29 if c1.size >= 29 else 0].astype(np.float32)   # points exactly on cell edges
    if c1.size > 40:
        c1.reshape(-1)[40] = np.nan
    h1 = voxel_hist(torch.from_numpy(c1).to(cuda)).cpu().numpy()
    h2 = voxel_hist(torch.from_numpy(c2).to(cuda)).cpu().numpy()
    # reference binning (utils.py:45-79 as restated in the product's numpy helper, itself pinned vs the oracle on CPU)
    r1 = u.get_voxel_occ_dist(c1, warning=False)
    r2 = u.get_voxel_occ_dist(c2, warning=False)
    assert h1.sum() > 0 and np.array_equal(h1 / h1.sum(), r1)
    assert np.array_equal(h2 / h2.sum(), r2)
    j = float(jsd_from_hists(torch.from_numpy(h1).to(cuda), torch.from_numpy(h2).to(cuda)))
    assert j == pytest.approx(u.JSD(c1, c2, warning=False), abs=1e-12)


def test_generation_metrics_fused_equals_reference_api(cuda, native_lib):
    from dpf_nets_b200.lib.networks import utils as u
    from dpf_nets_b200.lib.networks.evaluating import generation_metrics
    g = torch.Generator().manual_seed(3)
    a = (torch.rand((24, 256, 3), generator=g) - 0.5).to(cuda)
    b = (torch.rand((20, 256, 3), generator=g) * 0.8 - 0.4).to(cuda)
    r = generation_metrics(a, b)
    gg, tt, gt = u.pairwise_CD(a, a), u.pairwise_CD(b, b), u.pairwise_CD(a, b)
    assert r['COV-CD'] == pytest.approx(u.COV(gt), abs=1e-7)
    assert r['MMD-CD'] == pytest.approx(u.MMD(gt), rel=1e-6)
    assert r['1NN-CD'] == pytest.approx(u.KNN(gg, gt, tt, 1), abs=1e-7)
    assert r['JSD'] == pytest.approx(u.JSD(a.cpu().numpy(), b.cpu().numpy(), warning=False), abs=1e-9)
