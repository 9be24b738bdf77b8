"""`nn_distance(seta, setb) -> (dist1, dist2)`; same contract as the reference wrapper
(lib/metrics/pytorch_structural_losses/nn_distance.py:7-41)."""
from torch.autograd import Function

from .StructuralLossesBackend import NNDistance, NNDistanceGrad


class NNDistanceFunction(Function):
    @staticmethod
    def forward(ctx, seta, setb):
        dist1, idx1, dist2, idx2 = NNDistance(seta, setb)
        ctx.save_for_backward(seta, setb, idx1, idx2)
        return dist1, dist2

    @staticmethod
    def backward(ctx, grad_dist1, grad_dist2):
        seta, setb, idx1, idx2 = ctx.saved_tensors
        return tuple(NNDistanceGrad(seta, setb, idx1, idx2, grad_dist1, grad_dist2))


nn_distance = NNDistanceFunction.apply
