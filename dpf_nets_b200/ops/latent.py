"""Fused latent-side blocks (csrc/latent.cu): BatchNorm1d + Swish over (B,F) matrices and the RealNVPFlow transform, each
one kernel forward and one backward instead of the ~20 / ~45 launch-bound ATen kernels per coupling layer of the module
chains (reference: lib/networks/flows.py:163-213, encoders.py:31-83)."""
import ctypes

import torch

from .. import _lib
from . import _counters


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
        _counters.bump(bn.num_batches_tracked)
        momentum = bn.momentum if bn.momentum is not None else 1.0 / float(bn.num_batches_tracked + 1)
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


def _ptrs(ts):
    return (ctypes.c_void_p * len(ts))(*[None if t is None else t.data_ptr() for t in ts])


class _LatentFlowLayer(torch.autograd.Function):
    """One RealNVPFlow layer (csrc/latent_flow.cu): (g, [Wa, gamma, beta, Wb, bb] x {mu, logvar}) -> (g_out, mu, logvar)."""

    @staticmethod
    def forward(ctx, g, Wa0, ga0, be0, Wb0, bb0, Wa1, ga1, be1, Wb1, bb1, pos, keep_idx, rms, rvs, bn_eps, momentum, training, eps, inverse):
        g = g.contiguous()
        par = [t.detach().contiguous() for t in (Wa0, ga0, be0, Wb0, bb0, Wa1, ga1, be1, Wb1, bb1)]
        B, D = g.shape
        H, Kk = par[0].shape
        Wn = par[3].shape[0]
        dev = g.device
        g_out, mu, lv = torch.empty_like(g), torch.empty_like(g), torch.empty_like(g)
        hpre = torch.empty((2, B, H), dtype=torch.float32, device=dev)
        stat = torch.empty((2, 2, H), dtype=torch.float32, device=dev)
        raw = torch.empty((2, B, Wn), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.call("dpf_latent_flow_forward", g, pos, keep_idx, _ptrs([par[0], par[5]]), _ptrs([par[1], par[6]]), _ptrs([par[2], par[7]]),
                      _ptrs(rms), _ptrs(rvs), _ptrs([par[3], par[8]]), _ptrs([par[4], par[9]]), int(B), int(D), int(H), int(Kk), int(Wn),
                      ctypes.c_float(bn_eps), ctypes.c_float(momentum), bool(training), ctypes.c_float(eps), bool(inverse),
                      g_out, mu, lv, hpre, stat, raw, device=dev)
        ctx.save_for_backward(g, pos, keep_idx, hpre, stat, raw, *par)
        ctx.cfg = (bool(training), float(eps), bool(inverse))
        ctx.set_materialize_grads(False)
        return g_out, mu, lv

    @staticmethod
    def backward(ctx, dgo, dmu, dlv):
        g, pos, keep_idx, hpre, stat, raw, *par = ctx.saved_tensors
        training, eps, inverse = ctx.cfg
        B, D = g.shape
        H, Kk = par[0].shape
        Wn = par[3].shape[0]
        dev = g.device
        dg = torch.empty_like(g)
        grads = [torch.empty_like(t) for t in par]
        c = lambda t: None if t is None else t.contiguous()
        with torch.cuda.device(dev):
            _lib.call("dpf_latent_flow_backward", c(dgo), c(dmu), c(dlv), g, pos, keep_idx, _ptrs([par[0], par[5]]), _ptrs([par[1], par[6]]),
                      _ptrs([par[2], par[7]]), _ptrs([par[3], par[8]]), int(B), int(D), int(H), int(Kk), int(Wn), training,
                      ctypes.c_float(eps), inverse, hpre, stat, raw, dg, _ptrs([grads[0], grads[5]]), _ptrs([grads[1], grads[6]]),
                      _ptrs([grads[2], grads[7]]), _ptrs([grads[3], grads[8]]), _ptrs([grads[4], grads[9]]), device=dev)
        return (dg, *grads, None, None, None, None, None, None, None, None, None)


def latent_flow_layer_ok(g, nets):
    """the one-kernel layer handles B <= 64 rows, hidden / kept / warped widths % 32 == 0, affine BatchNorm with the same
    momentum in both branches."""
    (l0a, bn_a, l1a), (l0b, bn_b, l1b) = nets
    H, Kk = l0a.weight.shape
    Wn = l1a.weight.shape[0]
    if _lib.lib().dpf_latent_flow_supported(int(g.shape[0]), int(g.shape[1]), int(H), int(Kk), int(Wn)) != 0:
        return False
    return (l0a.bias is None and l0b.bias is None
            and l1a.bias is not None and l1b.bias is not None and bn_a.affine and bn_b.affine and bn_a.momentum == bn_b.momentum
            and bn_a.eps == bn_b.eps and bn_a.training == bn_b.training and bn_a.track_running_stats == bn_b.track_running_stats
            and (g.shape[0] > 1 or not bn_a.training))


def latent_flow_layer(g, nets, pos, keep_idx, eps, mode):
    """nets = ((Linear, BatchNorm1d, Linear) of the mu branch, the same of the logvar branch) -> (g_out, mu, logvar) with the
    modules' train / eval semantics (running-statistics update, num_batches_tracked)."""
    if mode not in ("direct", "inverse"):
        raise ValueError(mode)
    (l0a, bn_a, l1a), (l0b, bn_b, l1b) = nets
    training = bn_a.training or not bn_a.track_running_stats
    rms = rvs = [None, None]
    momentum = 0.0
    if bn_a.track_running_stats:
        if bn_a.training:
            _counters.bump(bn_a.num_batches_tracked)
            _counters.bump(bn_b.num_batches_tracked)
            momentum = bn_a.momentum if bn_a.momentum is not None else 1.0 / float(bn_a.num_batches_tracked + 1)
        rms, rvs = [bn_a.running_mean, bn_b.running_mean], [bn_a.running_var, bn_b.running_var]
    return _LatentFlowLayer.apply(g, l0a.weight, bn_a.weight, bn_a.bias, l1a.weight, l1a.bias, l0b.weight, bn_b.weight, bn_b.bias,
                                  l1b.weight, l1b.bias, pos, keep_idx, rms, rvs, float(bn_a.eps), float(momentum), training, float(eps),
                                  mode == "inverse")
