"""Point decoder = stack of 3*n_flows conditional coupling layers (reference:
lib/networks/decoders.py:41-72).  Output lists are indexed by layer - index 0 is always
flows[0].nvp1 - in both modes, exactly like the reference."""
from ._arena import CouplingStack
from ._flowfn import run_stack
from .flows import _triple_warps


class FlowOutputs(list):
    """A plain list of per-layer tensors that also carries the stacked (L,B,3,N) tensor it was
    unbound from and, for the logvar list, `total` = the sum over layers the kernels accumulated
    (what PointFlowNLL needs, losses.py:12-13) as its own autograd output."""
    stacked = None
    total = None


def _as_list(stacked, first=None, total=None):
    out = FlowOutputs(stacked.unbind(0))
    out.stacked = stacked
    if first is not None:
        # element 0 (samples[0] of the flow NLL) is the kernel's separate output Z == stacked[0]: its cotangent is one
        # (B,3,N) block instead of a zero-filled dense (L,B,3,N) gradient through unbind's backward
        out[0] = first
    out.total = total
    return out


class LocalCondRNVPDecoder(CouplingStack):
    def __init__(self, n_flows, f_n_features, g_n_features, weight_std=0.01):
        specs = []
        for i in range(n_flows):
            for j, w in enumerate(_triple_warps(i % 2)):
                specs.append(("flows.%d.nvp%d." % (i, j + 1), w))
        super().__init__(specs, f_n_features, g_n_features, weight_std=weight_std)
        self.n_flows = n_flows

    def forward(self, p, g, mode="direct"):
        P, MU, LV, Z, SLV = run_stack(self, p, g, mode)
        return _as_list(P, first=Z), _as_list(MU), _as_list(LV, total=SLV)

    def nll_terms(self, p, g, mode="inverse"):
        """-> (samples[0], sum over layers of logvar), each (B,3,N): everything PointFlowNLL reads from the
        decoder (losses.py:7-15), without materialising the per-layer mu outputs."""
        _, _, _, Z, SLV = run_stack(self, p, g, mode, want_mu=False)
        return Z, SLV


class TailedList(list):
    """[base] + decoder list: keeps a handle on the decoder's stacked tensor / layer total for fused reductions."""
    tail_stacked = None
    tail_total = None


def prepend(base, flow_list):
    out = TailedList([base] + list(flow_list))
    out.tail_stacked = getattr(flow_list, "stacked", None)
    out.tail_total = getattr(flow_list, "total", None)
    return out


import torch.nn as _nn

from .flows import RealNVPFlowCouple as _Couple


class GlobalRNVPDecoder(_nn.Module):
    """Latent prior flow: n_flows couples, list outputs ordered like the reference
    (lib/networks/decoders.py:7-38)."""

    def __init__(self, n_flows, n_features, g_n_features, weight_std=0.01):
        super().__init__()
        self.n_flows, self.n_features, self.g_n_features, self.weight_std = n_flows, n_features, g_n_features, weight_std
        self.flows = _nn.ModuleList([_Couple(n_features, g_n_features, weight_std=weight_std, pattern=i % 2)
                                     for i in range(n_flows)])

    def forward(self, g, mode='direct'):
        gs, mus, logvars = [], [], []
        cur = g
        order = range(self.n_flows) if mode == 'direct' else range(self.n_flows - 1, -1, -1)
        for i in order:
            o = self.flows[i](cur, mode=mode)
            if mode == 'direct':
                gs, mus, logvars = gs + o[0], mus + o[1], logvars + o[2]
                cur = gs[-1]
            else:
                gs, mus, logvars = o[0] + gs, o[1] + mus, o[2] + logvars
                cur = gs[0]
        return gs, mus, logvars
