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
