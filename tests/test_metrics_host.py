"""CPU: host-side metric helpers (COV/MMD/KNN/JSD) vs the numpy oracle, and the multi-rank sharding
logic (interleaved rows + one collective; gradient averaging) with world_size-2 gloo."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from oracle import metrics_oracle as mo
from oracle import structural as so

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _utils():
    from dpf_nets_b200.lib.networks import utils
    return utils


def test_cov_mmd_knn_jsd_vs_oracle():
    u = _utils()
    rng = np.random.default_rng(0)
    gg, tt, gt = [np.abs(rng.normal(size=(40, 40))).astype(np.float32) for _ in range(3)]
    gg = (gg + gg.T) / 2; np.fill_diagonal(gg, 0)
    tt = (tt + tt.T) / 2; np.fill_diagonal(tt, 0)
    T = torch.from_numpy
    assert u.COV(T(gt)) == pytest.approx(mo.cov(gt))
    assert u.MMD(T(gt)) == pytest.approx(mo.mmd(gt), rel=1e-6)
    assert u.KNN(T(gg), T(gt), T(tt), 1) == pytest.approx(mo.knn1(gg, gt, tt))
    c1 = rng.uniform(-0.5, 0.5, (6, 300, 3)).astype(np.float32)
    c2 = (rng.normal(size=(6, 300, 3)) * 0.2).astype(np.float32)
    assert u.JSD(c1, c2, warning=False) == pytest.approx(mo.jsd(c1, c2), abs=1e-9)
    assert u.JSD(c1, c1, warning=False) == pytest.approx(0.0, abs=1e-12)


def _worker(rank, world, port, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from dpf_nets_b200 import dist as dd
    rng = np.random.default_rng(3)
    A = rng.uniform(-0.5, 0.5, (7, 40, 3)).astype(np.float32)
    B = rng.uniform(-0.5, 0.5, (5, 33, 3)).astype(np.float32)
    full_ab = so.pairwise_cd(A, B)
    full_aa = so.pairwise_cd(A, A)

    def make(X, Y, full):
        def rows(out, row_start, row_step, n_rows, sym):
            for t in range(n_rows):
                i = row_start + t * row_step
                row = torch.from_numpy(full[i].copy())
                if sym:
                    row[:i] = 0
                out[i] = row
        return rows
    m1 = dd.sharded_pairwise(make(A, B, full_ab), 7, 5, torch.device("cpu"))
    m2 = dd.sharded_pairwise(make(A, A, full_aa), 7, 7, torch.device("cpu"), symmetric=True)
    lin = torch.nn.Linear(4, 3)
    torch.manual_seed(0)
    with torch.no_grad():
        for p in lin.parameters():
            p.copy_(torch.arange(p.numel(), dtype=torch.float32).view_as(p))
    x = torch.full((2, 4), float(rank + 1))
    lin(x).sum().backward()
    dd.allreduce_arena_grads(lin)
    q.put((rank, np.array_equal(m1.numpy(), full_ab), np.array_equal(m2.numpy(), full_aa),
           lin.weight.grad[0].tolist(), dd.shard_rows(7, rank, world)))
    dist.destroy_process_group()


def test_sharding_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 1000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    for rank, ok1, ok2, grad, shard in res:
        assert ok1 and ok2
        assert grad == pytest.approx([3.0] * 4)      # mean of 2*1 and 2*2
    assert res[0][4] == (0, 2, 4) and res[1][4] == (1, 2, 3)


# ---- multi-rank evaluate(): identical metrics at R = 1 and R = 2 (reference: evaluating.py:233-253) ----
class _FakeModel(torch.nn.Module):
    """Deterministic stand-in for the VAE in 'generating' / 'evaluating' mode: the 'generated' cloud is a
    fixed function of the input cloud, one of them NaN (exercises the duplicate substitution)."""

    def __init__(self, with_nan):
        super().__init__()
        self.w = torch.nn.Parameter(torch.zeros(1))
        self.with_nan = with_nan

    def forward(self, g_clouds, p_clouds, n_sampled_points=None):
        out = 0.9 * p_clouds.flip(2) + 0.01
        if self.with_nan:
            out = torch.where((p_clouds[:, :1, :1] > 0.31).expand_as(out), torch.full_like(out, float('nan')), out)
        return {'p_prior_samples': [out]}


def _patched_evaluate(util_mode, n_shapes=11, batch=4):
    """evaluate() on CPU with the CUDA compute replaced by the oracle (injected; the sharding / gathering /
    meter logic under test is the product's)."""
    from dpf_nets_b200 import dist as dd
    from dpf_nets_b200.lib.datasets.synthetic import SyntheticCloudDataset
    from dpf_nets_b200.lib.networks import evaluating as ev
    from torch.utils.data import DataLoader

    def cpu_pairwise(c1, c2, bs=2048):
        sym = c1 is c2
        full = so.pairwise_cd(c1.numpy(), c2.numpy())

        def rows(out, row_start, row_step, n_rows, s):
            for t in range(n_rows):
                i = row_start + t * row_step
                row = torch.from_numpy(full[i].copy())
                if s:
                    row[:i] = 0
                out[i] = row
        return dd.sharded_pairwise(rows, c1.shape[0], c2.shape[0], torch.device("cpu"), symmetric=sym)

    def cpu_scores(gg, gt, tt):
        return torch.tensor([mo.cov(gt.numpy()), mo.mmd(gt.numpy()), mo.knn1(gg.numpy(), gt.numpy(), tt.numpy())])

    def cpu_nnd(a, b):
        d1, _, d2, _ = so.nndistance(a.numpy(), b.numpy())
        return torch.from_numpy(d1), torch.from_numpy(d2)

    ev.pairwise_CD = cpu_pairwise
    ev._scores = cpu_scores
    ev._jsd = lambda g, r: torch.tensor(mo.jsd(g.numpy(), r.numpy()))
    ev.distChamferCUDA = cpu_nnd
    ds = SyntheticCloudDataset(n_shapes, cloud_size=48, part='test')
    rank, world = dd.world()
    sampler = dd.ShardSampler(len(ds), rank, world) if world > 1 else None
    it = DataLoader(ds, batch_size=batch, sampler=sampler, shuffle=False)
    return ev.evaluate(it, _FakeModel(util_mode == 'generating'), None, train_mode='p_rnvp_mc_g_rnvp_vae', util_mode=util_mode,
                       cloud_size=48, sampled_cloud_size=48, orig_scale_evaluation=False)


def _eval_worker(rank, world, port, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    res = {m: _patched_evaluate(m) for m in ("generating", "evaluating")}
    q.put((rank, res))
    dist.destroy_process_group()


def test_evaluate_world2_matches_world1():
    single = {m: _patched_evaluate(m) for m in ("generating", "evaluating")}
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 1000
    procs = [ctx.Process(target=_eval_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=180) for _ in range(2)), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
    for rank, r in res:
        for k in ("JSD", "COV-CD", "MMD-CD", "1NN-CD"):
            assert r["generating"][k] == pytest.approx(single["generating"][k], rel=1e-12, abs=1e-15), (rank, k)
        assert r["evaluating"]["CD"] == pytest.approx(single["evaluating"]["CD"], rel=1e-6)
    assert np.isfinite(single["generating"]["MMD-CD"]) and single["generating"]["COV-CD"] > 0


def test_shard_sampler_and_eval_scale():
    from dpf_nets_b200 import dist as dd
    from dpf_nets_b200.lib.networks.evaluating import _to_eval_scale
    assert [list(dd.ShardSampler(11, r, 3)) for r in range(3)] == [[0, 1, 2, 3], [4, 5, 6, 7], [8, 9, 10]]
    assert [len(dd.ShardSampler(2, r, 4)) for r in range(4)] == [1, 1, 0, 0]
    # the reference's in-place rescaling (evaluating.py:88-103): default generation config = cloud_scale 2.0, no
    # rescale2orig -> (x * 2 + 0) * orig_s + orig_c
    r = torch.ones(2, 3, 5); p = 2 * torch.ones(2, 3, 5)
    batch = {'orig_s': torch.tensor([2.0, 3.0]), 'orig_c': torch.tensor([[1., 0., 0.], [0., 1., 0.]])}
    kw = dict(orig_scale_evaluation=True, cloud_scale=True, cloud_scale_scale=2.0, cloud_translate=False,
              cloud_rescale2orig=False, cloud_recenter2orig=False)
    rr, pp = _to_eval_scale(r, p, batch, torch.device('cpu'), kw, 'generating')
    assert torch.equal(rr[0, :, 0], torch.tensor([5., 4., 4.])) and torch.equal(pp[1, :, 0], torch.tensor([12., 13., 12.]))
    # all_original config: clouds are already in the original frame -> only the constant shift is undone
    kw2 = dict(orig_scale_evaluation=True, cloud_scale=False, cloud_translate=True, cloud_translate_shift=[0.5, 0., -0.5],
               cloud_rescale2orig=True, cloud_recenter2orig=True)
    rr, _ = _to_eval_scale(r, p, {}, torch.device('cpu'), kw2, 'evaluating')
    assert torch.equal(rr[0, :, 0], torch.tensor([1.5, 1.0, 0.5]))
    with pytest.raises(KeyError):
        _to_eval_scale(r, p, {}, torch.device('cpu'), kw, 'generating')
    assert _to_eval_scale(r, p, {}, torch.device('cpu'), dict(orig_scale_evaluation=False), 'generating')[0] is r
