"""Flow NLL / entropy losses with the reference's class names and formulas
(lib/networks/losses.py:7-51).  The sum over the per-layer log-dets uses the stacked tensor
carried by the decoder's output lists when present: the kernels' own per-point sum over layers (`total`),
else one reduction over the stacked tensor, instead of 63 full-tensor adds."""
import math

import torch
import torch.nn as nn


def _sum_over_layers(logvars):
    """logvars[0] + ... + logvars[-1] like the reference's Python sum()."""
    if len(logvars) > 1:
        tail = logvars[1:] if isinstance(logvars, list) else list(logvars)[1:]
        stacked = getattr(tail, "stacked", None)
        total = getattr(logvars, "tail_total", None)
        if total is not None:            # accumulated per point in the decoder kernels' epilogues
            return logvars[0] + total
        src = getattr(logvars, "tail_stacked", None)
        if src is not None:
            return logvars[0] + src.sum(0)
        if stacked is not None:
            return logvars[0] + stacked.sum(0)
    if len(logvars) > 2 and all(lv.shape == logvars[1].shape for lv in logvars[2:]) and logvars[1].is_cuda:
        # latent flows: 14 (B,G) tensors - one stack + one reduction instead of a chain of launch-bound adds
        return logvars[0] + torch.stack(list(logvars[1:])).sum(0)
    acc = logvars[0]
    for lv in logvars[1:]:
        acc = acc + lv
    return acc


class PointFlowNLL(nn.Module):
    def forward(self, samples, mus, logvars):
        s0 = samples[0]
        quad = (s0 - mus[0]) ** 2 / torch.exp(logvars[0])
        return 0.5 * (torch.sum(_sum_over_layers(logvars) + quad) / s0.shape[0]
                      + math.log(2.0 * math.pi) * s0.shape[1] * s0.shape[2])


class GaussianFlowNLL(nn.Module):
    def forward(self, samples, mus, logvars):
        s0 = samples[0]
        quad = (s0 - mus[0]) ** 2 / torch.exp(logvars[0])
        return 0.5 * (torch.sum(_sum_over_layers(logvars) + quad) / s0.shape[0]
                      + math.log(2.0 * math.pi) * s0.shape[1])


class GaussianEntropy(nn.Module):
    def forward(self, logvars):
        return 0.5 * (logvars.shape[1] * (1.0 + math.log(2.0 * math.pi)) + logvars.sum(1).mean())


class Local_Cond_RNVP_MC_Global_RNVP_VAE_Loss(nn.Module):
    def __init__(self, **kwargs):
        super().__init__()
        self.pnll_weight = kwargs.get('pnll_weight')
        self.gnll_weight = kwargs.get('gnll_weight')
        self.gent_weight = kwargs.get('gent_weight')
        self.PNLL = PointFlowNLL()
        self.GNLL = GaussianFlowNLL()
        self.GENT = GaussianEntropy()

    def forward(self, g_clouds, l_clouds, outputs):
        pnll = self.PNLL(outputs['p_prior_samples'], outputs['p_prior_mus'], outputs['p_prior_logvars'])
        gnll = self.GNLL(outputs['g_prior_samples'], outputs['g_prior_mus'], outputs['g_prior_logvars'])
        gent = self.GENT(outputs['g_posterior_logvars'])
        return self.pnll_weight * pnll + self.gnll_weight * gnll - self.gent_weight * gent, pnll, gnll, gent
