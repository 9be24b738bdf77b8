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
