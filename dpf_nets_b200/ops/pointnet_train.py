"""Train-mode PointNet cloud encoder + max-pool as ONE autograd function over the library's own kernels
(reference: PointNetCloudEncoder.features in .train(), lib/networks/encoders.py:9-28, + torch.max over dim 2,
lib/networks/models.py:130-131, + torch.autograd through all of it).

Forward (csrc/pointnet_layers.cu, csrc/pointnet_train.cu):
    layer 0 (3 -> 64)     never materialised: its batch statistics are analytic in the moments of the input cloud
                          (mean_c = W0[c] . E[x], var_c = W0[c]^T Cov(x) W0[c]); A_0 = relu(a . x + c) is recomputed on load
    layer 1 (64 -> 128)   Z_1 = W1 A_0      tcgen05 kernel, output statistics in the epilogue
    layer 2 (128 -> 256)  Z_2 = W2 A_1      same kernel, A_1 = relu(sc_1 Z_1 + sh_1) applied on load
    layer 3 (256 -> 512)  statistics + max / min over the points of W3 A_2 (pool kernel, A_2 applied on load; never stored)
Only Z_1 and Z_2 are stored (100 MB at 32 x 2048; the module chain keeps 12 activations, 600 MB).

Backward: the pooled layer's analytic sparse backward (ops/pointnet_pool.py), then per layer one streaming reduction
(BatchNorm-backward sums), one dgrad kernel and one wgrad kernel whose operand loaders apply the BatchNorm + ReLU backward
to (dA, Z) on the fly; layer 0's parameter gradients follow from four sums per channel and the input moments."""
import ctypes

import torch

from .. import _lib
from . import pointnet_pool as _pp
from .pointnet_pool import pool_backward, pool_select

LD_X3, LD_AFFINE, LD_BNBWD = 0, 1, 2
WIDTHS = (3, 64, 128, 256, 512)


def _image(W, R, K, transpose):
    """bf16 hi | lo swizzled image of the [R x K] matrix W (or W^T: element (r, k) = W[k, r])."""
    nb = ctypes.c_longlong(0)
    _lib.check(_lib.lib().dpf_pointnet_layer_image_bytes(int(R), int(K), ctypes.byref(nb)), "dpf_pointnet_layer_image_bytes")
    img = torch.empty(int(nb.value), dtype=torch.uint8, device=W.device)
    rs, cs = (1, W.shape[1]) if transpose else (W.shape[1], 1)
    _lib.call("dpf_pointnet_layer_pack", W, int(R), int(K), ctypes.c_longlong(rs), ctypes.c_longlong(cs), img, device=W.device)
    return img


def _groups(B, N):
    g = ctypes.c_int(0)
    _lib.check(_lib.lib().dpf_pointnet_layer_groups(int(B), int(N), ctypes.byref(g)), "dpf_pointnet_layer_groups")
    return int(g.value)


def _merge_stats(stat):
    """(G, C, 3) {count, mean, M2} per work item -> (mean (C,), biased variance (C,)) in double (Chan et al.)."""
    n, m, m2 = stat[..., 0].double(), stat[..., 1].double(), stat[..., 2].double()
    tot = n.sum(0)
    mean = (n * m).sum(0) / tot
    var = (m2.sum(0) + (n * (m - mean) ** 2).sum(0)) / tot
    return mean, var


def _table(*cols, width=8):
    """per-channel loader table (C, 8) fp32 from column vectors (missing columns zero)."""
    C = cols[0].shape[0]
    t = torch.zeros((C, width), dtype=torch.float32, device=cols[0].device)
    for j, c in enumerate(cols):
        if c is not None:
            t[:, j] = c.to(torch.float32)
    return t


def _gemm(loader, K, in0, in1, tab, img, B, N, Mout, want_out=True, row_off=None, want_stat=False):
    dev = in0.device
    out = torch.empty((B, Mout, N), dtype=torch.float32, device=dev) if want_out else None
    stat = torch.empty((B * _groups(B, N), Mout, 3), dtype=torch.float32, device=dev) if want_stat else None
    _lib.call("dpf_pointnet_layer_gemm", int(loader), int(K), in0, in1, tab, img, int(B), int(N), int(Mout), out, row_off, stat, device=dev)
    return out, stat


class PointNetTrainFunction(torch.autograd.Function):
    """(x (B,3,N), W0..W3, gamma0..3, beta0..3, eps) -> (out (B,512), [mean_l, var_l for l = 0..3] (biased batch statistics))."""

    @staticmethod
    def forward(ctx, x, W0, W1, W2, W3, g0, g1, g2, g3, b0, b1, b2, b3, eps):
        x = x.contiguous()
        _lib.require_cuda(x)
        B, _, N = x.shape
        M = B * N
        dev = x.device
        W0, W1, W2, W3 = (w.detach().contiguous() for w in (W0, W1, W2, W3))
        # ---- layer 0: analytic batch statistics from the moments of the cloud (double) ----
        xd = x.double()
        xm = xd.mean((0, 2))
        xc = xd - xm.view(1, 3, 1)
        cov = torch.einsum('bin,bjn->ij', xc, xc) / M
        W0d = W0.double()
        mean0 = W0d @ xm
        var0 = torch.einsum('ci,ij,cj->c', W0d, cov, W0d).clamp_min(0.0)
        istd0 = torch.rsqrt(var0 + eps)
        sc0 = g0.detach().double() * istd0
        tab0 = _table(sc0 * W0d[:, 0], sc0 * W0d[:, 1], sc0 * W0d[:, 2], b0.detach().double() - sc0 * mean0)
        # ---- layer 1 ----
        with torch.cuda.device(dev):
            Z1, st1 = _gemm(LD_X3, 64, x, None, tab0, _image(W1, 128, 64, False), B, N, 128, want_stat=True)
            mean1, var1 = _merge_stats(st1)
            istd1 = torch.rsqrt(var1 + eps)
            sc1 = g1.detach().double() * istd1
            tab1 = _table(sc1, b1.detach().double() - sc1 * mean1, None, None, None, mean1)
            # ---- layer 2 ----
            Z2, st2 = _gemm(LD_AFFINE, 128, Z1, None, tab1, _image(W2, 256, 128, False), B, N, 256, want_stat=True)
            mean2, var2 = _merge_stats(st2)
            istd2 = torch.rsqrt(var2 + eps)
            sc2 = g2.detach().double() * istd2
            tab2 = _table(sc2, b2.detach().double() - sc2 * mean2, None, None, None, mean2)
            # ---- layer 3 + max-pool ----
            mean3, var3, vmax, vmin, imax, imin = _pp._pool_stats(Z2, W3, in_tab=tab2)
        out, sel = pool_select(g3.detach(), b3.detach(), mean3, var3, vmax, vmin, imax, imin, eps, x.dtype)
        ctx.save_for_backward(x, W0, W1, W2, W3, g0, g1, g2, g3, Z1, Z2, tab0, tab1, tab2, xm, cov, istd0, istd1, istd2, *sel)
        ctx.eps = eps
        stats = [t.to(x.dtype) for t in (mean0, var0, mean1, var1, mean2, var2, mean3, var3)]
        ctx.mark_non_differentiable(*stats)
        return (out, *stats)

    @staticmethod
    def backward(ctx, dout, *_unused):
        (x, W0, W1, W2, W3, g0, g1, g2, g3, Z1, Z2, tab0, tab1, tab2, xm, cov, istd0, istd1, istd2, *sel) = ctx.saved_tensors
        B, _, N = x.shape
        M = B * N
        dev = x.device
        with torch.cuda.device(dev):
            # ---- layer 3 + max-pool: analytic sparse backward on the recomputed activation A_2 ----
            A2 = torch.relu_(torch.addcmul(tab2[:, 1].view(1, -1, 1), Z2, tab2[:, 0].view(1, -1, 1)))
            dA2, dW3, dg3, db3 = pool_backward(A2, W3, g3, sel, dout.contiguous(), need_input=True, need_weight=True)
            del A2
            # ---- layer 2 ----
            dA1, dW2, dg2, db2 = _layer_backward(dA2, Z2, tab2, g2, istd2, W2, 256, 128, M, B, N, LD_AFFINE, Z1, tab1)
            del dA2
            # ---- layer 1 ----
            dA0, dW1, dg1, db1 = _layer_backward(dA1, Z1, tab1, g1, istd1, W1, 128, 64, M, B, N, LD_X3, x, tab0)
            del dA1
            # ---- layer 0: four sums per channel + the input moments ----
            sums = torch.zeros((64, 4), dtype=torch.float64, device=dev)
            _lib.call("dpf_pointnet_layer0_bwd_sums", dA0, x, tab0, int(B), 64, int(N), sums, device=dev)
        S, T = sums[:, 0], sums[:, 1:4]
        W0d = W0.double()
        Tc = T - S.unsqueeze(1) * xm.unsqueeze(0)                    # sum dy (x_j - mean_j)
        db0 = S
        dg0 = istd0 * (W0d * Tc).sum(1)
        gi = g0.double() * istd0
        dW0 = gi.unsqueeze(1) * (Tc - (dg0 / M * istd0).unsqueeze(1) * (M * (W0d @ cov)))
        dt = x.dtype
        return (None, dW0.to(dt), dW1, dW2, dW3, dg0.to(dt), dg1, dg2, dg3, db0.to(dt), db1, db2, db3, None)


def _layer_backward(dA, Z, tab, gamma, istd, W, C, Cprev, M, B, N, q_loader, q_in, q_tab):
    """BatchNorm + ReLU + SharedDot backward of one layer: (dA (B,C,N), Z (B,C,N)) -> (dA_prev (B,Cprev,N), dW (C,Cprev), dgamma, dbeta)."""
    dev = dA.device
    sums = torch.zeros((C, 2), dtype=torch.float64, device=dev)
    _lib.call("dpf_pointnet_bn_bwd_sums", dA, Z, tab, int(B), int(C), int(N), sums, device=dev)
    dbeta = sums[:, 0]
    dgamma = sums[:, 1] * istd
    g = gamma.detach().double() * istd
    tabb = tab.clone()
    tabb[:, 2] = g.float()
    tabb[:, 3] = (g * dbeta / M).float()
    tabb[:, 4] = (g * dgamma / M * istd).float()
    dA_prev, _ = _gemm(LD_BNBWD, C, dA, Z, tabb, _image(W, Cprev, C, True), B, N, Cprev)
    dW = torch.zeros((C, Cprev), dtype=torch.float32, device=dev)
    _lib.call("dpf_pointnet_layer_wgrad", int(C), int(Cprev), 0, int(q_loader), dA, Z, tabb, q_in, q_tab, int(B), int(N), dW, device=dev)
    return dA_prev, dW, dgamma.to(dA.dtype), dbeta.to(dA.dtype)


def pointnet_train_forward(x, sds, bns):
    """x (B,3,N); sds / bns: the four SharedDot / BatchNorm1d modules of PointNetCloudEncoder.features -> (B,512);
    updates the running statistics of the four BatchNorm modules like the module chain would."""
    B, _, N = x.shape
    Ws = [sd.weight[0] for sd in sds]
    res = PointNetTrainFunction.apply(x, *Ws, *[bn.weight for bn in bns], *[bn.bias for bn in bns], float(bns[0].eps))
    out, stats = res[0], res[1:]
    with torch.no_grad():
        M = B * N
        for l, bn in enumerate(bns):
            if not bn.track_running_stats:
                continue
            bn.num_batches_tracked += 1
            mom = bn.momentum if bn.momentum is not None else 1.0 / float(bn.num_batches_tracked)
            bn.running_mean.mul_(1 - mom).add_(stats[2 * l], alpha=mom)
            bn.running_var.mul_(1 - mom).add_(stats[2 * l + 1] * (M / max(M - 1, 1)), alpha=mom)
    return out
