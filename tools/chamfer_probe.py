"""Quick GPU probe: timing of our Chamfer kernels (the comparison with the reference kernel lives in tests/test_chamfer_gpu.py)."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dpf_nets_b200.ops import pairwise_cd  # noqa: E402
from dpf_nets_b200.lib.metrics.StructuralLosses import StructuralLossesBackend as B  # noqa: E402


def ev_time(fn, reps=3, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e-3)
    return min(ts)


def main():
    S = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    N = 2048
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(0)
    A = (torch.rand((S, N, 3), generator=g) - 0.5).to(dev)
    Bc = (torch.rand((S, N, 3), generator=g) - 0.5).to(dev)
    res = {"S": S, "N": N}
    t = ev_time(lambda: pairwise_cd(A, Bc))
    res["pairwise_cd_s"] = t
    res["pairs_per_s"] = S * S / t
    res["point_pair_evals_per_s"] = S * S * N * N / t
    t = ev_time(lambda: pairwise_cd(A, A, symmetric=True))
    res["pairwise_cd_sym_s"] = t
    t = ev_time(lambda: B.NNDistance(A, Bc))
    res["nndistance_b%d_s" % S] = t
    res["nndistance_pairs_per_s"] = S / t
    print(json.dumps(res))


if __name__ == "__main__":
    main()
