"""Train-mode last layer of the PointNet cloud encoder fused with the max-pool over the points
(reference: features.{sd2, sd2_bn, sd2_relu} in .train(), lib/networks/encoders.py:9-28, + torch.max over dim 2,
lib/networks/models.py:130-131).

Forward: ONE pass of a tcgen05 kernel over h2 (dpf_pointnet_pool_forward) yields the batch statistics of h = W h2 and,
per (shape, channel), max / min of h with their point indices; BatchNorm + ReLU are monotone per channel, so
    out[b,c] = relu(gamma_c (h* - mu_c) / sigma_c + beta_c),   h* = max_n h (gamma_c >= 0) or min_n h (gamma_c < 0).
The (B,512,N) activation is never materialised.

Backward: the max-pool gradient is SPARSE (one point per (shape, channel)) and BatchNorm's batch terms are LINEAR in h:
    dh[p,c] = (1/sigma_c) (gamma_c d'[b,c] [p = n*(b,c)] - a1_c - xhat[p,c] a2_c),     a1 = gamma dbeta / M,  a2 = gamma dgamma / M
so with S = sum_p h2[p], G = sum_p h2[p] h2[p]^T (256 x 256 Gram matrix) and C = W^T diag(a2 / sigma^2) W
    dW  = diag(gamma/sigma) T  -  (a1/sigma) S^T  -  diag(a2/sigma^2) W Gc,        T[c] = sum_b d'[b,c] h2[b,:,n*(b,c)]
    dh2 = scatter(coef[b,c] W[c,:] at point n*(b,c))  -  W^T(a1/sigma)  -  C (h2 - m)
(m = S / M, Gc = sum_p (h2[p] - m)(h2[p] - m)^T: the centred forms avoid the cancellation of W G - mu S^T in fp32)
i.e. two 256 x 256 GEMMs over the points (library GEMMs) + gathers / scatters of B x 512 rows, instead of the library path's
512 x 256 dgrad + wgrad and three passes over the 134 MB activation."""
import ctypes

import torch

from .. import _lib
from . import _counters


def pool_stats(h2, W, in_tab=None, want_asum=False, merge=True):
    """h2 (B,256,N), W (512,256) CUDA fp32 -> (mean (512,), biased var (512,) of h = W a over all B*N points (double), vmax,
    vmin (B,512), imax, imin (B,512) i64[, asum (B,256): sum of a over the points]); a = h2, or relu(sc h2 + sh) with the
    per-channel table in_tab (256, 8) {sc, sh, ..}."""
    _lib.require_cuda(h2, W, in_tab)
    B, Cin, N = h2.shape
    if Cin != 256 or tuple(W.shape) != (512, 256) or h2.dtype != torch.float32 or W.dtype != torch.float32:
        raise _lib.DpfNativeError("pointnet pool kernel is specialised on a 256 -> 512 last layer (got %s, %s)"
                                  % (tuple(h2.shape), tuple(W.shape)))
    dev = h2.device
    nb = ctypes.c_longlong(0)
    _lib.check(_lib.lib().dpf_pointnet_pool_workspace_bytes(ctypes.byref(nb)), "dpf_pointnet_pool_workspace_bytes")
    ws = torch.empty(int(nb.value), dtype=torch.uint8, device=dev)
    stat = torch.empty((B, 512, 2), dtype=torch.float32, device=dev)
    vmax = torch.empty((B, 512), dtype=torch.float32, device=dev)
    vmin = torch.empty((B, 512), dtype=torch.float32, device=dev)
    imax = torch.empty((B, 512), dtype=torch.int32, device=dev)
    imin = torch.empty((B, 512), dtype=torch.int32, device=dev)
    asum = torch.empty((B, 256), dtype=torch.float32, device=dev) if want_asum else None
    with torch.cuda.device(dev):
        _lib.call("dpf_pointnet_pool_forward_ex", h2, in_tab, W, int(B), int(N), ws, stat, vmax, vmin, imax, imin, asum, device=dev)
    if not merge:       # raw per-shape {mean, M2} (B,512,2): the caller merges them (dpf_pointnet_stats_finalize)
        res = (stat, vmax, vmin, imax.long(), imin.long())
        return res + (asum,) if want_asum else res
    # merge the B equal-sized groups (count N, mean, M2) into the batch statistics (Chan et al.), in double
    gm, gm2 = stat[..., 0].double(), stat[..., 1].double()
    mean = gm.mean(0)
    var = (gm2.sum(0) + N * ((gm - mean) ** 2).sum(0)) / float(B * N)
    res = (mean, var, vmax, vmin, imax.long(), imin.long())
    return res + (asum,) if want_asum else res


_pool_stats = pool_stats


def pool_select(gamma, beta, mean64, var64, vmax, vmin, imax, imin, eps, dtype):
    """BatchNorm + ReLU + max over the points from the per-(shape, channel) extrema -> (out (B,512), saved tensors
    (mean, sigma, xhat, idx, y) for pool_backward)."""
    mean, var = mean64.to(dtype), var64.clamp_min(0.0).to(dtype)
    sigma = torch.sqrt(var + eps)
    pos = gamma >= 0
    hsel = torch.where(pos.unsqueeze(0), vmax.to(dtype), vmin.to(dtype))
    idx = torch.where(pos.unsqueeze(0), imax, imin)
    xhat = (hsel - mean) / sigma
    y = xhat * gamma + beta
    return torch.relu(y), (mean, sigma, xhat, idx, y)


def pool_backward(h2, W, gamma, sel, dout, need_input=True, need_weight=True):
    """-> (dh2 (B,256,N) | None, dW (512,256) | None, dgamma, dbeta); h2 = the layer's (post-ReLU) input activations."""
    mean, sigma, xhat, idx, y = sel
    B, Cin, N = h2.shape
    M = B * N
    d = dout * (y > 0).to(dout.dtype)                       # (B,C): cotangent of y at the selected point
    dbeta = d.sum(0)
    dgamma = (d * xhat).sum(0)
    a1 = gamma * dbeta / M
    a2 = gamma * dgamma / M
    inv = 1.0 / sigma
    coef = d * (gamma * inv)                                 # (B,C): dh at the selected point (before the batch terms)
    dW = dh2 = None
    if need_input or need_weight:
        gidx = idx.unsqueeze(1).expand(B, Cin, idx.shape[1])        # (B,Cin,C)
        S = h2.sum((0, 2))                                                       # (Cin,)
        # CENTRED activations: h - mu = W (h2 - m) exactly, so the batch terms are formed from deviations and not from
        # differences of large sums (W G - mu S^T cancels catastrophically in fp32 over 65 536 points)
        hc = h2 - (S / M).view(1, Cin, 1)
    if need_weight:
        Gc = torch.matmul(hc, hc.transpose(1, 2)).sum(0)                         # (Cin,Cin) covariance-form Gram matrix
        hsel_rows = torch.gather(h2, 2, gidx)                                    # (B,Cin,C): h2 at the selected points
        T = torch.einsum('bc,bkc->ck', coef, hsel_rows)
        dW = T - (a1 * inv).unsqueeze(1) * S.unsqueeze(0) - (a2 * inv * inv).unsqueeze(1) * torch.matmul(W, Gc)
    if need_input:
        w_scaled = W * (a2 * inv * inv).unsqueeze(1)                             # diag(a2/sigma^2) W
        Cmat = torch.matmul(W.t(), w_scaled)                                     # (Cin,Cin)
        const = torch.mv(W.t(), a1 * inv)                                        # (Cin,)
        dh2 = -const.view(1, Cin, 1) - torch.matmul(Cmat, hc)
        contrib = coef.unsqueeze(1) * W.t().unsqueeze(0)                         # (B,Cin,C): coef[b,c] W[c,k]
        dh2.scatter_add_(2, gidx, contrib)
    return dh2, dW, dgamma, dbeta


class PooledLastLayer(torch.autograd.Function):
    """(h2 (B,256,N), W (512,256), gamma, beta (512,)) -> (out (B,512), batch mean (512,), biased batch var (512,))."""

    @staticmethod
    def forward(ctx, h2, W, gamma, beta, eps):
        h2 = h2.contiguous()
        W = W.contiguous()
        mean64, var64, vmax, vmin, imax, imin = _pool_stats(h2.detach(), W.detach())     # module attribute: tests substitute it
        out, sel = pool_select(gamma, beta, mean64, var64, vmax, vmin, imax, imin, eps, h2.dtype)
        ctx.save_for_backward(h2, W, gamma, *sel)
        mean, var = sel[0], var64.clamp_min(0.0).to(h2.dtype)
        ctx.mark_non_differentiable(mean, var)
        return out, mean, var

    @staticmethod
    def backward(ctx, dout, _dmean, _dvar):
        h2, W, gamma, *sel = ctx.saved_tensors
        dh2, dW, dgamma, dbeta = pool_backward(h2, W, gamma, sel, dout, ctx.needs_input_grad[0], ctx.needs_input_grad[1])
        return dh2, dW, dgamma, dbeta, None


def pooled_bn_relu_max(h2, weight, bn):
    """Train-mode  max_n relu(bn(weight @ h2))  with nn.BatchNorm1d `bn`'s parameters; updates its running statistics
    (momentum, unbiased variance) and num_batches_tracked like the module would."""
    B, _, N = h2.shape
    out, mean, var = PooledLastLayer.apply(h2, weight, bn.weight, bn.bias, float(bn.eps))
    if bn.track_running_stats:
        with torch.no_grad():
            M = B * N
            _counters.bump(bn.num_batches_tracked)
            mom = bn.momentum if bn.momentum is not None else 1.0 / float(bn.num_batches_tracked + 1)
            bn.running_mean.mul_(1 - mom).add_(mean, alpha=mom)
            bn.running_var.mul_(1 - mom).add_(var * (M / max(M - 1, 1)), alpha=mom)
    return out
