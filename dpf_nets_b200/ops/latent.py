"""Fused latent-side blocks (csrc/latent.cu): BatchNorm1d + Swish over (B,F) matrices and the RealNVPFlow transform, each
one kernel forward and one backward instead of the ~20 / ~45 launch-bound ATen kernels per coupling layer of the module
chains (reference: lib/networks/flows.py:163-213, encoders.py:31-83)."""
import ctypes

import torch

from .. import _lib


def fused_ok(*tensors):
    return all(t.is_cuda and t.dtype == torch.float32 for t in tensors)


class _BnSwish(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, gamma, beta, rm, rv, eps, momentum, training):
        x = x.contiguous()
        B, F = x.shape
        y = torch.empty_like(x)
        save_mean = torch.empty(F, dtype=torch.float32, device=x.device)
        save_istd = torch.empty(F, dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            _lib.call("dpf_bn_swish_forward", x, gamma.detach().contiguous(), beta.detach().contiguous(), rm, rv, int(B), int(F),
                      ctypes.c_float(eps), ctypes.c_float(momentum), bool(training), y, save_mean, save_istd, device=x.device)
        ctx.save_for_backward(x, gamma, beta, save_mean, save_istd)
        ctx.training = bool(training)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, gamma, beta, save_mean, save_istd = ctx.saved_tensors
        B, F = x.shape
        dx = torch.empty_like(x)
        dgamma = torch.empty_like(save_mean)
        dbeta = torch.empty_like(save_mean)
        with torch.cuda.device(x.device):
            _lib.call("dpf_bn_swish_backward", dy.contiguous(), x, gamma.detach().contiguous(), beta.detach().contiguous(), save_mean,
                      save_istd, int(B), int(F), ctx.training, dx, dgamma, dbeta, device=x.device)
        return dx, dgamma, dbeta, None, None, None, None, None


def bn_swish(x, bn):
    """swish(bn(x)) for an nn.BatchNorm1d `bn` (affine, tracking running statistics) on a (B,F) CUDA fp32 matrix, with the
    module's train / eval semantics (running-statistics update, num_batches_tracked)."""
    training = bn.training or not bn.track_running_stats
    rm = bn.running_mean if bn.track_running_stats else None
    rv = bn.running_var if bn.track_running_stats else None
    momentum = 0.0
    if bn.training and bn.track_running_stats:
        bn.num_batches_tracked += 1
        momentum = bn.momentum if bn.momentum is not None else 1.0 / float(bn.num_batches_tracked)
    elif training:
        rm = rv = None
    return _BnSwish.apply(x, bn.weight, bn.bias, rm, rv, float(bn.eps), float(momentum), training)


class _LatentAffine(torch.autograd.Function):
    @staticmethod
    def forward(ctx, g, raw_mu, raw_lv, pos, eps, inverse):
        g, raw_mu, raw_lv = g.contiguous(), raw_mu.contiguous(), raw_lv.contiguous()
        B, G = g.shape
        W = raw_mu.shape[1]
        g_out, mu, lv = torch.empty_like(g), torch.empty_like(g), torch.empty_like(g)
        with torch.cuda.device(g.device):
            _lib.call("dpf_latent_affine_forward", g, raw_mu, raw_lv, pos, int(B), int(G), int(W), ctypes.c_float(eps), bool(inverse),
                      g_out, mu, lv, device=g.device)
        ctx.save_for_backward(g, raw_mu, raw_lv, pos)
        ctx.eps, ctx.inverse = eps, bool(inverse)
        ctx.set_materialize_grads(False)
        return g_out, mu, lv

    @staticmethod
    def backward(ctx, dgo, dmu, dlv):
        g, raw_mu, raw_lv, pos = ctx.saved_tensors
        B, G = g.shape
        W = raw_mu.shape[1]
        dg, draw_mu, draw_lv = torch.empty_like(g), torch.empty_like(raw_mu), torch.empty_like(raw_lv)
        c = lambda t: None if t is None else t.contiguous()
        with torch.cuda.device(g.device):
            _lib.call("dpf_latent_affine_backward", c(dgo), c(dmu), c(dlv), g, raw_mu, raw_lv, pos, int(B), int(G), int(W),
                      ctypes.c_float(ctx.eps), ctx.inverse, dg, draw_mu, draw_lv, device=g.device)
        return dg, draw_mu, draw_lv, None, None, None


def latent_affine(g, raw_mu, raw_lv, pos, eps, mode):
    """-> (g_out, mu, logvar), each (B,G); pos (G,) int32 CUDA: index into the warp list or -1."""
    if mode not in ("direct", "inverse"):
        raise ValueError(mode)
    return _LatentAffine.apply(g, raw_mu, raw_lv, pos, float(eps), mode == "inverse")
