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
from . import _counters
from . import pointnet_pool as _pp
from .pointnet_pool import pool_select

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


def _groups(B, N, Mout):
    g = ctypes.c_int(0)
    _lib.check(_lib.lib().dpf_pointnet_layer_groups(int(B), int(N), int(Mout), ctypes.byref(g)), "dpf_pointnet_layer_groups")
    return int(g.value)


def _merge_stats(stat):
    """(G, C, 3) {count, mean, M2} per work item -> (mean (C,), biased variance (C,)) in double (Chan et al.)."""
    n, m, m2 = stat[..., 0].double(), stat[..., 1].double(), stat[..., 2].double()
    tot = n.sum(0)
    mean = (n * m).sum(0) / tot
    var = (m2.sum(0) + (n * (m - mean) ** 2).sum(0)) / tot
    return mean, var


def _table(*cols, width=8):
    """per-channel loader table (C, 8) fp32 from column vectors (missing columns zero); one stack + one cast."""
    ref = cols[0]
    zero = torch.zeros_like(ref)
    cols = [zero if c is None else c.to(ref.dtype) for c in cols] + [zero] * (width - len(cols))
    return torch.stack(cols, 1).to(torch.float32).contiguous()


def _wgrad(MP, NQ, gram, q_loader, p_in0, p_in1, p_tab, q_in, q_tab, B, N):
    """(MP, NQ) = sum over all points of P Q^T (pl_wgrad_kernel + the reduction of its per-CTA partial sums)."""
    dev = p_in0.device
    nb = ctypes.c_longlong(0)
    _lib.check(_lib.lib().dpf_pointnet_layer_wgrad_scratch_bytes(int(MP), int(NQ), int(B), int(N), ctypes.byref(nb)), "dpf_pointnet_layer_wgrad_scratch_bytes")
    scratch = torch.empty(int(nb.value), dtype=torch.uint8, device=dev)
    out = torch.zeros((MP, NQ), dtype=torch.float32, device=dev)
    _lib.call("dpf_pointnet_layer_wgrad", int(MP), int(NQ), int(gram), int(q_loader), p_in0, p_in1, p_tab, q_in, q_tab, int(B), int(N), scratch, out, device=dev)
    return out


def _gemm(loader, K, in0, in1, tab, img, B, N, Mout, want_out=True, row_off=None, want_stat=False):
    dev = in0.device
    out = torch.empty((B, Mout, N), dtype=torch.float32, device=dev) if want_out else None
    stat = torch.empty((_groups(B, N, Mout), Mout, 3), dtype=torch.float32, device=dev) if want_stat else None
    _lib.call("dpf_pointnet_layer_gemm", int(loader), int(K), in0, in1, tab, img, int(B), int(N), int(Mout), out, row_off, stat, device=dev)
    return out, stat


def _finalize_stats(stat, width, count, gamma, beta, eps, want_tab=True):
    """per-CTA statistics -> (loader-1 table (C,8) | None, stats (C,3) {mean, biased var, istd}); one kernel."""
    G, C = stat.shape[0], stat.shape[1]
    dev = stat.device
    tab = torch.empty((C, 8), dtype=torch.float32, device=dev) if want_tab else None
    st = torch.empty((C, 3), dtype=torch.float32, device=dev)
    _lib.call("dpf_pointnet_stats_finalize", stat, int(G), int(C), int(width), float(count), gamma if want_tab else None,
              beta if want_tab else None, float(eps), tab, st, device=dev)
    return tab, st


class PointNetTrainFunction(torch.autograd.Function):
    """(x (B,3,N), W0..W3, gamma0..3, beta0..3, eps) -> (out (B,512), st0..st3: (C_l, 3) {batch mean, biased batch var, istd})."""

    @staticmethod
    def forward(ctx, x, W0, W1, W2, W3, g0, g1, g2, g3, b0, b1, b2, b3, eps):
        x = x.contiguous()
        _lib.require_cuda(x)
        B, _, N = x.shape
        dev = x.device
        W0, W1, W2, W3 = (w.detach().contiguous() for w in (W0, W1, W2, W3))
        g0, g1, g2, g3, b0, b1, b2, b3 = (t.detach().contiguous() for t in (g0, g1, g2, g3, b0, b1, b2, b3))
        with torch.cuda.device(dev):
            # ---- layer 0: analytic batch statistics from the moments of the cloud ----
            moments = torch.zeros(9, dtype=torch.float64, device=dev)
            _lib.call("dpf_pointnet_input_moments", x, int(B), int(N), moments, device=dev)
            tab0 = torch.empty((64, 8), dtype=torch.float32, device=dev)
            st0 = torch.empty((64, 3), dtype=torch.float32, device=dev)
            _lib.call("dpf_pointnet_layer0_finalize", moments, W0, g0, b0, 64, int(B), int(N), float(eps), tab0, st0, device=dev)
            # ---- layer 1 ----
            Z1, stat1 = _gemm(LD_X3, 64, x, None, tab0, _image(W1, 128, 64, False), B, N, 128, want_stat=True)
            tab1, st1 = _finalize_stats(stat1, 3, 0.0, g1, b1, eps)
            # ---- layer 2 ----
            Z2, stat2 = _gemm(LD_AFFINE, 128, Z1, None, tab1, _image(W2, 256, 128, False), B, N, 256, want_stat=True)
            tab2, st2 = _finalize_stats(stat2, 3, 0.0, g2, b2, eps)
            # ---- layer 3 + max-pool ----
            stat3, vmax, vmin, imax, imin, asum = _pp._pool_stats(Z2, W3, in_tab=tab2, want_asum=True, merge=False)
            _, st3 = _finalize_stats(stat3, 2, float(N), None, None, eps, want_tab=False)
        out, sel = pool_select(g3, b3, st3[:, 0], st3[:, 1], vmax, vmin, imax, imin, eps, x.dtype)
        ctx.save_for_backward(x, W0, W1, W2, W3, g0, g1, g2, g3, Z1, Z2, tab0, tab1, tab2, moments, st0, st1, st2, asum, *sel)
        ctx.mark_non_differentiable(st0, st1, st2, st3)
        return out, st0, st1, st2, st3

    @staticmethod
    def backward(ctx, dout, *_unused):
        (x, W0, W1, W2, W3, g0, g1, g2, g3, Z1, Z2, tab0, tab1, tab2, moments, st0, st1, st2, asum, *sel) = ctx.saved_tensors
        B, _, N = x.shape
        dev = x.device
        with torch.cuda.device(dev):
            # ---- layer 3 + max-pool: analytic sparse backward, A_2 recomputed inside the kernels' loaders ----
            dA2, dW3, dg3, db3 = _pool_layer_backward(Z2, tab2, asum, W3, g3, sel, dout.contiguous(), B, N)
            # ---- layer 2 ----
            dA1, dW2, dg2, db2 = _layer_backward(dA2, Z2, tab2, g2, st2, W2, 256, 128, B, N, LD_AFFINE, Z1, tab1)
            del dA2
            # ---- layer 1 ----
            dA0, dW1, dg1, db1 = _layer_backward(dA1, Z1, tab1, g1, st1, W1, 128, 64, B, N, LD_X3, x, tab0)
            del dA1
            # ---- layer 0: four sums per channel + the input moments ----
            sums = torch.zeros((64, 4), dtype=torch.float64, device=dev)
            _lib.call("dpf_pointnet_layer0_bwd_sums", dA0, x, tab0, int(B), 64, int(N), sums, device=dev)
            dW0 = torch.empty((64, 3), dtype=torch.float32, device=dev)
            dg0 = torch.empty(64, dtype=torch.float32, device=dev)
            db0 = torch.empty(64, dtype=torch.float32, device=dev)
            _lib.call("dpf_pointnet_layer0_bwd_finalize", sums, moments, W0, g0, st0, 64, int(B), int(N), dW0, dg0, db0, device=dev)
        return (None, dW0, dW1, dW2, dW3, dg0, dg1, dg2, dg3, db0, db1, db2, db3, None)


def _sparse_backward(Z2, tab2, idx, coef, W, B, N, T, dA2):
    """T += sum_b coef a(n*), dA2[.., n*] += coef W (csrc/pointnet_layers.cu::pl_pool_sparse_bwd_kernel)."""
    _lib.call("dpf_pointnet_pool_sparse_backward", Z2, tab2, idx.contiguous(), coef, W, int(B), int(N), int(W.shape[0]), T, dA2, device=Z2.device)


def _pool_layer_backward(Z2, tab2, asum, W, gamma, sel, dout, B, N):
    """ops/pointnet_pool.py::pool_backward with its two 256 x 256 GEMMs over the points on the library's own kernels: the Gram
    matrix of the centred activations (pl_wgrad_kernel, gram form) and -C (A_2 - m) - const (pl_gemm_kernel with a row
    offset); A_2 = relu(sc Z_2 + sh) and the centring are applied by the operand loaders, the sums S come from the forward."""
    mean, sigma, xhat, idx, y = sel
    dev = Z2.device
    M = B * N
    Cin = 256
    d = dout * (y > 0).to(dout.dtype)
    dbeta = d.sum(0)
    dgamma = (d * xhat).sum(0)
    a1 = gamma * dbeta / M
    a2 = gamma * dgamma / M
    inv = 1.0 / sigma
    coef = d * (gamma * inv)
    S = asum.double().sum(0)
    sc, sh = tab2[:, 0], tab2[:, 1]
    tabc = _table(sc, sh, (S / M).to(sc.dtype))
    Gc = _wgrad(Cin, Cin, 1, LD_AFFINE, Z2, None, tabc, None, None, B, N)
    w_scaled = W * (a2 * inv * inv).unsqueeze(1)
    negC = -torch.matmul(W.t(), w_scaled)                                           # -(W^T diag(a2 / sigma^2) W)
    negconst = -torch.mv(W.t(), a1 * inv)
    dA2, _ = _gemm(LD_AFFINE, Cin, Z2, None, tabc, _image(negC.contiguous(), Cin, Cin, False), B, N, Cin, row_off=negconst.contiguous())
    # the sparse part (one selected point per (shape, channel)): T = sum_b coef a(n*) and dA2[.., n*] += coef W, one kernel
    T = torch.zeros((W.shape[0], Cin), dtype=Z2.dtype, device=dev)
    _sparse_backward(Z2, tab2, idx, coef.contiguous(), W, B, N, T, dA2)
    Sf = S.to(Z2.dtype)
    dW = T - (a1 * inv).unsqueeze(1) * Sf.unsqueeze(0) - (a2 * inv * inv).unsqueeze(1) * torch.matmul(W, Gc)
    return dA2, dW, dgamma, dbeta


def _layer_backward(dA, Z, tab, gamma, st, W, C, Cprev, B, N, q_loader, q_in, q_tab):
    """BatchNorm + ReLU + SharedDot backward of one layer: (dA (B,C,N), Z (B,C,N)) -> (dA_prev (B,Cprev,N), dW (C,Cprev), dgamma, dbeta)."""
    dev = dA.device
    sums = torch.zeros((C, 2), dtype=torch.float64, device=dev)
    _lib.call("dpf_pointnet_bn_bwd_sums", dA, Z, tab, int(B), int(C), int(N), sums, device=dev)
    tabb = torch.empty((C, 8), dtype=torch.float32, device=dev)
    dgamma = torch.empty(C, dtype=torch.float32, device=dev)
    dbeta = torch.empty(C, dtype=torch.float32, device=dev)
    _lib.call("dpf_pointnet_bwd_finalize", sums, tab, gamma, st, int(C), int(B), int(N), tabb, dgamma, dbeta, device=dev)
    dA_prev, _ = _gemm(LD_BNBWD, C, dA, Z, tabb, _image(W, Cprev, C, True), B, N, Cprev)
    dW = _wgrad(C, Cprev, 0, q_loader, dA, Z, tabb, q_in, q_tab, B, N)
    return dA_prev, dW, dgamma, dbeta


def pointnet_train_forward(x, sds, bns):
    """x (B,3,N); sds / bns: the four SharedDot / BatchNorm1d modules of PointNetCloudEncoder.features -> (B,512);
    updates the running statistics of the four BatchNorm modules like the module chain would."""
    B, _, N = x.shape
    Ws = [sd.weight[0] for sd in sds]
    res = PointNetTrainFunction.apply(x, *Ws, *[bn.weight for bn in bns], *[bn.bias for bn in bns], float(bns[0].eps))
    out, stats = res[0], res[1:]
    with torch.no_grad():
        M = B * N
        for st, bn in zip(stats, bns):
            if not bn.track_running_stats:
                continue
            _counters.bump(bn.num_batches_tracked)
            mom = bn.momentum if bn.momentum is not None else 1.0 / float(bn.num_batches_tracked + 1)
            bn.running_mean.mul_(1 - mom).add_(st[:, 0], alpha=mom)
            bn.running_var.mul_(1 - mom).add_(st[:, 1], alpha=mom * (M / max(M - 1, 1)))
    return out
