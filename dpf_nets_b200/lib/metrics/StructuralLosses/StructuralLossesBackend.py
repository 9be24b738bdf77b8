"""Same five callables as the reference's pybind module `StructuralLossesBackend`
(lib/metrics/pytorch_structural_losses/pybind/bind.cpp:9-15, src/structural_loss.cpp:22-124),
implemented over the C ABI of libdpfnets_b200.so.  Inputs must be CUDA, contiguous, float32;
outputs are allocated here on the inputs' device; kernels run on torch's current stream of
THAT device (the reference launches on the current device regardless of the tensors')."""
import torch

from .... import _lib


def _check(*ts):
    for t in ts:
        if not t.is_cuda:
            raise RuntimeError("must be a CUDA tensor")
        if not t.is_contiguous():
            raise RuntimeError("must be contiguous")
        if t.dtype not in (torch.float32, torch.int32):
            raise RuntimeError("must be float32 / int32")


def NNDistance(set_d, set_q):
    _check(set_d, set_q)
    b, n, m = set_d.size(0), set_d.size(1), set_q.size(1)
    dev = set_d.device
    dist1 = torch.empty((b, n), dtype=torch.float32, device=dev)
    idx1 = torch.empty((b, n), dtype=torch.int32, device=dev)
    dist2 = torch.empty((b, m), dtype=torch.float32, device=dev)
    idx2 = torch.empty((b, m), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _lib.call("dpf_nndistance", b, n, set_d, m, set_q, dist1, idx1, dist2, idx2, device=dev)
    return [dist1, idx1, dist2, idx2]


def NNDistanceGrad(set_d, set_q, idx1, idx2, grad_dist1, grad_dist2):
    grad_dist1 = grad_dist1.contiguous()
    grad_dist2 = grad_dist2.contiguous()
    _check(set_d, set_q, idx1, idx2, grad_dist1, grad_dist2)
    b, n, m = set_d.size(0), set_d.size(1), set_q.size(1)
    dev = set_d.device
    grad1 = torch.empty((b, n, 3), dtype=torch.float32, device=dev)
    grad2 = torch.empty((b, m, 3), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.call("dpf_nndistance_grad", b, n, set_d, m, set_q, grad_dist1, idx1, grad_dist2, idx2,
                  grad1, grad2, device=dev)
    return [grad1, grad2]


def ApproxMatch(set_d, set_q):
    """-> [match (b, m, n), temp (b, 2(n+m))]  (structural_loss.cpp:22-37; temp is unused scratch)."""
    _check(set_d, set_q)
    b, n, m = set_d.size(0), set_d.size(1), set_q.size(1)
    dev = set_d.device
    match = torch.empty((b, m, n), dtype=torch.float32, device=dev)
    temp = torch.empty((b, (n + m) * 2), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.call("dpf_approxmatch", b, n, m, set_d, set_q, match, temp, device=dev)
    return [match, temp]


def MatchCost(set_d, set_q, match):
    _check(set_d, set_q, match)
    b, n, m = set_d.size(0), set_d.size(1), set_q.size(1)
    out = torch.empty((b,), dtype=torch.float32, device=set_d.device)
    with torch.cuda.device(set_d.device):
        _lib.call("dpf_matchcost", b, n, m, set_d, set_q, match, out, device=set_d.device)
    return out


def MatchCostGrad(set_d, set_q, match):
    _check(set_d, set_q, match)
    b, n, m = set_d.size(0), set_d.size(1), set_q.size(1)
    dev = set_d.device
    grad1 = torch.empty((b, n, 3), dtype=torch.float32, device=dev)
    grad2 = torch.empty((b, m, 3), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.call("dpf_matchcost_grad", b, n, m, set_d, set_q, match, grad1, grad2, device=dev)
    return [grad1, grad2]
