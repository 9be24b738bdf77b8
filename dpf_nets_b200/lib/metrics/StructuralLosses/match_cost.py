"""`match_cost(seta, setb) -> cost (b,)`; same contract as the reference wrapper
(lib/metrics/pytorch_structural_losses/match_cost.py:6-44)."""
from torch.autograd import Function

from .StructuralLossesBackend import ApproxMatch, MatchCost, MatchCostGrad


class MatchCostFunction(Function):
    @staticmethod
    def forward(ctx, seta, setb):
        match, _ = ApproxMatch(seta, setb)
        ctx.save_for_backward(seta, setb, match)
        return MatchCost(seta, setb, match)

    @staticmethod
    def backward(ctx, grad_output):
        seta, setb, match = ctx.saved_tensors
        grada, gradb = MatchCostGrad(seta, setb, match)
        scale = grad_output.unsqueeze(1).unsqueeze(2)
        return grada * scale, gradb * scale


match_cost = MatchCostFunction.apply
