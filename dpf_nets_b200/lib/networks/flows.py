"""Conditional affine-coupling flow layers with the reference's class names, constructor
arguments, forward contract and state_dict keys (lib/networks/flows.py:10-160); the per-point
conditioner, scale/shift and log-det all run in the fused sm_100a kernels.

mode='direct'  : p_out = sqrt(eps + exp(logvar)) * p + mu      (noise -> data, sampling)
mode='inverse' : p_out = (p - mu) / sqrt(eps + exp(logvar))    (data -> noise, training NLL)
"""
from ._arena import CouplingStack
from ._flowfn import run_stack


def _triple_warps(pattern):
    if pattern == 0:
        return [[0], [1], [2]]
    if pattern == 1:
        return [[0, 1], [0, 2], [1, 2]]
    raise ValueError("pattern must be 0 or 1")


class CondRealNVPFlow3D(CouplingStack):
    """One coupling layer: forward(p (B,3,N), g (B,G), mode) -> (p_out, mu, logvar)."""

    def __init__(self, f_n_features, g_n_features, weight_std=0.01, warp_inds=[0],
                 centered_translation=False, eps=1e-6):
        super().__init__([("", list(warp_inds))], f_n_features, g_n_features, weight_std=weight_std, eps=eps)
        self.warp_inds = list(warp_inds)
        self.keep_inds = [c for c in (0, 1, 2) if c not in self.warp_inds]
        self.centered_translation = centered_translation  # accepted and ignored, like the reference

    def forward(self, p, g, mode="direct"):
        P, MU, LV, _, _ = run_stack(self, p, g, mode)
        return P[0], MU[0], LV[0]


class CondRealNVPFlow3DTriple(CouplingStack):
    """Three coupling layers nvp1..nvp3; returns lists ordered [1, 2, 3] in both modes."""

    def __init__(self, f_n_features, g_n_features, weight_std=0.02, pattern=0, centered_translation=False):
        specs = [("nvp%d." % (j + 1), w) for j, w in enumerate(_triple_warps(pattern))]
        super().__init__(specs, f_n_features, g_n_features, weight_std=weight_std)
        self.pattern = pattern
        self.centered_translation = centered_translation

    def forward(self, p, g, mode="direct"):
        P, MU, LV, _, _ = run_stack(self, p, g, mode)
        return list(P.unbind(0)), list(MU.unbind(0)), list(LV.unbind(0))


# ---- latent-space (per-shape) coupling flows: (B, G) vectors, launch-latency only -------------
import numpy as _np
import torch as _torch
import torch.nn as _nn

from .layers import Swish as _Swish


class RealNVPFlow(_nn.Module):
    """Unconditional coupling layer on the shape latent (reference flows.py:163-213):
    logvar = log(eps + exp(net(g_keep))), g_out = exp(+-logvar/2) * g (+-) mu."""

    def __init__(self, n_features, g_n_features, weight_std=0.01, warp_inds=[0], eps=1e-6):
        super().__init__()
        self.n_features, self.g_n_features, self.weight_std = n_features, g_n_features, weight_std
        self.warp_inds = [int(i) for i in warp_inds]
        self.keep_inds = [i for i in range(g_n_features) if i not in set(self.warp_inds)]
        self.register_buffer('eps', _torch.tensor([eps], dtype=_torch.float32))
        # device-resident index vectors (not in the state_dict): indexing with Python lists would build a CPU
        # index tensor and copy it to the GPU on every call - a sync per layer, and illegal under graph capture
        self.register_buffer('_keep_idx', _torch.tensor(self.keep_inds, dtype=_torch.long), persistent=False)
        self.register_buffer('_warp_idx', _torch.tensor(self.warp_inds, dtype=_torch.long), persistent=False)
        self.register_buffer('_keep_idx32', _torch.tensor(self.keep_inds, dtype=_torch.int32), persistent=False)
        pos = _np.full(g_n_features, -1, dtype=_np.int32)
        pos[_np.asarray(self.warp_inds, dtype=_np.int64)] = _np.arange(len(self.warp_inds), dtype=_np.int32)
        self.register_buffer('_pos', _torch.from_numpy(pos), persistent=False)     # latent position -> index in the warp list / -1
        self.eps_value = float(_np.float32(eps))
        for br in ('mu', 'logvar'):
            net = _nn.Sequential()
            net.add_module(br + '_mlp0', _nn.Linear(len(self.keep_inds), n_features, bias=False))
            net.add_module(br + '_mlp0_bn', _nn.BatchNorm1d(n_features))
            net.add_module(br + '_mlp0_swish', _Swish())
            net.add_module(br + '_mlp1', _nn.Linear(n_features, len(self.warp_inds), bias=True))
            with _torch.no_grad():
                net[-1].weight.normal_(std=weight_std)
                net[-1].bias.zero_()
            setattr(self, 'T_%s_0' % br, net)

    fused = True      # class switch: False keeps the plain module chain on the GPU too (tests compare both)
    fused_layer = True      # class switch: False keeps the per-block fused kernels (BatchNorm + Swish, transform) with library GEMMs

    def _net(self, net, br, kept):
        """Linear -> BatchNorm1d + Swish (one fused kernel) -> Linear of one branch."""
        from ...ops.latent import bn_swish
        h = getattr(net, br + '_mlp0')(kept)
        return getattr(net, br + '_mlp1')(bn_swish(h, getattr(net, br + '_mlp0_bn')))

    def forward(self, g, mode='direct'):
        if self.fused and g.is_cuda and g.dtype == _torch.float32 and g.dim() == 2:
            # fused path (csrc/latent.cu): BatchNorm + Swish and the whole transform are one kernel each, forward and backward
            from ...ops.latent import latent_affine, latent_flow_layer, latent_flow_layer_ok
            nets = tuple((getattr(n, b + '_mlp0'), getattr(n, b + '_mlp0_bn'), getattr(n, b + '_mlp1'))
                         for n, b in ((self.T_mu_0, 'mu'), (self.T_logvar_0, 'logvar')))
            if self.fused_layer and latent_flow_layer_ok(g, nets):
                # the whole layer, forward and backward, as one kernel each (csrc/latent_flow.cu)
                return latent_flow_layer(g, nets, self._pos, self._keep_idx32, self.eps_value, mode)
            kept = g.index_select(1, self._keep_idx)
            raw_lv = self._net(self.T_logvar_0, 'logvar', kept)
            raw_mu = self._net(self.T_mu_0, 'mu', kept)
            return latent_affine(g, raw_mu, raw_lv, self._pos, self.eps_value, mode)
        kept = g.index_select(1, self._keep_idx)
        logvar = _torch.zeros_like(g).index_copy(1, self._warp_idx, _torch.log(self.eps + _torch.exp(self.T_logvar_0(kept))))
        mu = _torch.zeros_like(g).index_copy(1, self._warp_idx, self.T_mu_0(kept))
        if mode == 'direct':
            g_out = _torch.exp(0.5 * logvar) * g + mu
        elif mode == 'inverse':
            g_out = _torch.exp(-0.5 * logvar) * (g - mu)
        else:
            raise ValueError(mode)
        return g_out, mu, logvar


class RealNVPFlowCouple(_nn.Module):
    """Two complementary latent coupling layers: even/odd (pattern 0) or halves (pattern 1)."""

    def __init__(self, n_features, g_n_features, weight_std=0.01, pattern=0):
        super().__init__()
        idx = _np.arange(g_n_features)
        parts = (idx[::2], idx[1::2]) if pattern == 0 else (idx[:g_n_features // 2], idx[g_n_features // 2:])
        self.pattern = pattern
        self.nvp1 = RealNVPFlow(n_features, g_n_features, weight_std=weight_std, warp_inds=list(parts[0]))
        self.nvp2 = RealNVPFlow(n_features, g_n_features, weight_std=weight_std, warp_inds=list(parts[1]))

    def forward(self, g, mode='direct'):
        if mode == 'direct':
            a = self.nvp1(g, mode=mode)
            b = self.nvp2(a[0], mode=mode)
        else:
            b = self.nvp2(g, mode=mode)
            a = self.nvp1(b[0], mode=mode)
        return [a[0], b[0]], [a[1], b[1]], [a[2], b[2]]
