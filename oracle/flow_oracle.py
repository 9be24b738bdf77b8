"""TEST INFRASTRUCTURE ONLY - CPU fp32 restatement of the reference's point-flow decoder math.

Explicit tensor formulas (no nn.Module) for:
  * SharedDot                     lib/networks/layers.py:13-45
  * CondRealNVPFlow3D             lib/networks/flows.py:10-117     (`coupling_forward`)
  * CondRealNVPFlow3DTriple       lib/networks/flows.py:120-160    (`triple_warps`)
  * LocalCondRNVPDecoder          lib/networks/decoders.py:41-72   (`decoder_forward`)
  * PointFlowNLL                  lib/networks/losses.py:7-15      (`point_flow_nll`)
  * hand-derived backward of one coupling layer - `coupling_backward`
    (checked against torch.autograd of `coupling_forward` and of the reference module in tests/).

Parameters are addressed by the reference's own state_dict key names relative to one
CondRealNVPFlow3D module (e.g. 'T_mu_0.mu_sd0.weight').  Pinned by tests/golden/*.pt, produced by
importing the reference from /root/reference (tests/golden/make_golden.py).
"""
import math

import torch

BN_EPS = 1e-5
BN_MOMENTUM = 0.1
BRANCHES = ("mu", "logvar")
# When True the 64x64 SharedDot of the conditioner is evaluated like the BF16 tensor-core path
# (bf16 operands, fp32 accumulate): the checker for the kernel's bf16 mode on ill-conditioned
# fixtures, where plain fp32-vs-bf16 differences are dominated by BatchNorm noise amplification.
EMULATE_BF16_GEMM = False


def triple_warps(pattern):
    """warp_inds of nvp1..nvp3 (flows.py:130-148)."""
    return [[0], [1], [2]] if pattern == 0 else [[0, 1], [0, 2], [1, 2]]


def decoder_layer_names(n_flows):
    """(key prefix, warp_inds) of the 3*n_flows coupling layers in list order (decoders.py:49-52)."""
    out = []
    for i in range(n_flows):
        for j, w in enumerate(triple_warps(i % 2)):
            out.append(("flows.%d.nvp%d." % (i, j + 1), w))
    return out


def _bn(x, w, b, rm, rv, training, dims):
    """BatchNorm1d (eps 1e-5, momentum 0.1): returns y, (mean, biased var) used, new running stats."""
    if training:
        n = 1
        for d in dims:
            n *= x.shape[d]
        mean = x.mean(dim=dims)
        var = x.var(dim=dims, unbiased=False)
        new_rm = (1 - BN_MOMENTUM) * rm + BN_MOMENTUM * mean
        new_rv = (1 - BN_MOMENTUM) * rv + BN_MOMENTUM * var * (n / max(n - 1, 1))
    else:
        mean, var, new_rm, new_rv = rm, rv, rm, rv
    shape = [1] * x.dim()
    shape[1] = -1
    y = (x - mean.view(shape)) / torch.sqrt(var.view(shape) + BN_EPS)
    if w is not None:
        y = y * w.view(shape) + b.view(shape)
    return y, mean, var, new_rm, new_rv


def film_net(P, br, kind, g, training, new_stats=None):
    """T_{br}_0_cond_{kind}: Linear(G->F, no bias) . BN(batch) . Swish . Linear(F->F)  (flows.py:33-45)."""
    pre = "T_%s_0_cond_%s.%s_sd1_film_%s" % (br, kind, br, kind)
    u = g @ P[pre + "0.weight"].t()
    y, _, _, rm, rv = _bn(u, P[pre + "0_bn.weight"], P[pre + "0_bn.bias"], P[pre + "0_bn.running_mean"],
                          P[pre + "0_bn.running_var"], training, (0,))
    if new_stats is not None:
        new_stats[pre + "0_bn.running_mean"] = rm
        new_stats[pre + "0_bn.running_var"] = rv
    y = y * torch.sigmoid(y)
    return y @ P[pre + "1.weight"].t() + P[pre + "1.bias"]


def conditioner(P, br, xk, g, training, new_stats=None, keep=None):
    """One branch (mu or logvar) of the per-point conditioner; returns o (B,w,N) and intermediates."""
    eps = P["eps"]
    t0 = "T_%s_0.%s_" % (br, br)
    h0 = torch.matmul(P[t0 + "sd0.weight"][0], xk)                     # (B,F,N)  SharedDot, no bias
    z, mean_a, var_a, rm, rv = _bn(h0, P[t0 + "sd0_bn.weight"], P[t0 + "sd0_bn.bias"],
                                   P[t0 + "sd0_bn.running_mean"], P[t0 + "sd0_bn.running_var"], training, (0, 2))
    if new_stats is not None:
        new_stats[t0 + "sd0_bn.running_mean"], new_stats[t0 + "sd0_bn.running_var"] = rm, rv
    h1 = torch.relu(z)
    if EMULATE_BF16_GEMM:   # operands of the 64x64 SharedDot rounded to bf16, fp32 accumulation
        h2pre = torch.matmul(P[t0 + "sd1.weight"][0].bfloat16().to(h1.dtype), h1.bfloat16().to(h1.dtype))
    else:
        h2pre = torch.matmul(P[t0 + "sd1.weight"][0], h1)
    h2n, mean_b, var_b, rm, rv = _bn(h2pre, None, None, P[t0 + "sd1_bn.running_mean"],
                                     P[t0 + "sd1_bn.running_var"], training, (0, 2))
    if new_stats is not None:
        new_stats[t0 + "sd1_bn.running_mean"], new_stats[t0 + "sd1_bn.running_var"] = rm, rv
    wraw = film_net(P, br, "w", g, training, new_stats)
    s = eps + torch.exp(wraw)                                           # (B,F)
    t = film_net(P, br, "b", g, training, new_stats)
    a = s.unsqueeze(2) * h2n + t.unsqueeze(2)
    h3 = torch.relu(a)
    t1 = "T_%s_1.%s_sd2." % (br, br)
    o = torch.matmul(P[t1 + "weight"][0], h3) + P[t1 + "bias"][0].view(1, -1, 1)
    inter = dict(h0=h0, z=z, h1=h1, h2pre=h2pre, h2n=h2n, a=a, h3=h3, s=s, t=t, wraw=wraw,
                 mean_a=mean_a, var_a=var_a, mean_b=mean_b, var_b=var_b)
    return o, inter


def coupling_forward(P, p, g, mode, warp_inds, training=False, new_stats=None, return_inter=False):
    """CondRealNVPFlow3D.forward (flows.py:95-117).  p (B,3,N), g (B,G) -> p_out, mu, logvar."""
    keep = [c for c in (0, 1, 2) if c not in warp_inds]
    eps = P["eps"]
    xk = p[:, keep, :].contiguous()
    logvar = torch.zeros_like(p)
    mu = torch.zeros_like(p)
    o_lv, i_lv = conditioner(P, "logvar", xk, g, training, new_stats)
    logvar[:, warp_inds, :] = o_lv / (1 + o_lv.abs())                   # softsign
    o_mu, i_mu = conditioner(P, "mu", xk, g, training, new_stats)
    mu[:, warp_inds, :] = o_mu
    sigma = torch.sqrt(eps + torch.exp(logvar))
    if mode == "direct":
        p_out = sigma * p + mu
    elif mode == "inverse":
        p_out = (p - mu) / sigma
    else:
        raise ValueError(mode)
    if return_inter:
        return p_out, mu, logvar, dict(mu=i_mu, logvar=i_lv, o_mu=o_mu, o_logvar=o_lv, sigma=sigma, keep=keep)
    return p_out, mu, logvar


def decoder_forward(layers, p, g, mode, training=False, new_stats=None):
    """LocalCondRNVPDecoder.forward (decoders.py:54-72).  `layers` = list of (param dict, warp_inds)
    in list order [flows[0].nvp1, flows[0].nvp2, ...]; returns lists indexed like the reference."""
    L = len(layers)
    ps, mus, lvs = [None] * L, [None] * L, [None] * L
    order = range(L) if mode == "direct" else range(L - 1, -1, -1)
    cur = p
    for l in order:
        P, warp = layers[l]
        ns = None
        if new_stats is not None:
            ns = {}
        cur, mus[l], lvs[l] = coupling_forward(P, cur, g, mode, warp, training, ns)
        if new_stats is not None:
            new_stats[l] = ns
        ps[l] = cur
    return ps, mus, lvs


def point_flow_nll(samples, mus, logvars):
    """PointFlowNLL (losses.py:7-15)."""
    s0 = samples[0]
    acc = logvars[0]
    for lv in logvars[1:]:
        acc = acc + lv
    return 0.5 * (torch.sum(acc + (s0 - mus[0]) ** 2 / torch.exp(logvars[0])) / s0.shape[0]
                  + math.log(2.0 * math.pi) * s0.shape[1] * s0.shape[2])


# ------------------------------------------------------------------------------------------------
# Hand-derived backward of one coupling layer in TRAIN mode (batch-stat BN), the formulas the
# fused CUDA backward implements.  Returns grads keyed like the parameters + dp + dg.
# ------------------------------------------------------------------------------------------------
def _bn_train_backward(dy, xhat, istd, dims):
    n = 1
    for d in dims:
        n *= dy.shape[d]
    shape = [1] * dy.dim()
    shape[1] = -1
    m1 = dy.sum(dim=dims) / n
    m2 = (dy * xhat).sum(dim=dims) / n
    return istd.view(shape) * (dy - m1.view(shape) - xhat * m2.view(shape))


def film_net_backward(P, br, kind, g, dout, training, grads):
    """Backward of film_net; accumulates parameter grads into `grads`, returns dg."""
    pre = "T_%s_0_cond_%s.%s_sd1_film_%s" % (br, kind, br, kind)
    W0, W1 = P[pre + "0.weight"], P[pre + "1.weight"]
    u = g @ W0.t()
    if training:
        mean, var = u.mean(0), u.var(0, unbiased=False)
    else:
        mean, var = P[pre + "0_bn.running_mean"], P[pre + "0_bn.running_var"]
    istd = 1.0 / torch.sqrt(var + BN_EPS)
    xhat = (u - mean) * istd
    gam, bet = P[pre + "0_bn.weight"], P[pre + "0_bn.bias"]
    y = xhat * gam + bet
    sg = torch.sigmoid(y)
    sw = y * sg
    grads[pre + "1.weight"] = dout.t() @ sw
    grads[pre + "1.bias"] = dout.sum(0)
    dsw = dout @ W1
    dy = dsw * (sg + y * sg * (1 - sg))
    grads[pre + "0_bn.weight"] = (dy * xhat).sum(0)
    grads[pre + "0_bn.bias"] = dy.sum(0)
    dxhat = dy * gam
    if training:
        du = _bn_train_backward(dxhat, xhat, istd, (0,))
    else:
        du = dxhat * istd
    grads[pre + "0.weight"] = du.t() @ g
    return du @ W0


def coupling_backward(P, p, g, mode, warp_inds, dy, dmu_ext=None, dlv_ext=None, training=True):
    """Gradients of a scalar loss given dy = dL/dp_out, dmu_ext = dL/dmu, dlv_ext = dL/dlogvar
    (each (B,3,N) or None).  Returns (dp, dg, grads dict)."""
    eps = P["eps"]
    p_out, mu, logvar, I = coupling_forward(P, p, g, mode, warp_inds, training, None, return_inter=True)
    keep = I["keep"]
    B, _, N = p.shape
    sigma = I["sigma"]
    grads = {}
    dp = torch.zeros_like(p)
    if mode == "inverse":
        dp[:] = dy / sigma                                  # kept channels: sigma = sqrt(eps+1)
        dmu = -dy / sigma
        dl = -dy * p_out * torch.exp(logvar) / (2 * sigma * sigma)
    else:
        dp[:] = dy * sigma
        dmu = dy.clone()
        dl = dy * p * torch.exp(logvar) / (2 * sigma)
    if dmu_ext is not None:
        dmu = dmu + dmu_ext
    if dlv_ext is not None:
        dl = dl + dlv_ext
    do = {"mu": dmu[:, warp_inds, :], "logvar": dl[:, warp_inds, :] / (1 + I["o_logvar"].abs()) ** 2}
    xk = p[:, keep, :]
    dg = torch.zeros_like(g)
    for br in BRANCHES:
        it = I[br]
        t0 = "T_%s_0.%s_" % (br, br)
        t1 = "T_%s_1.%s_sd2." % (br, br)
        W0, W1, W2 = P[t0 + "sd0.weight"][0], P[t0 + "sd1.weight"][0], P[t1 + "weight"][0]
        d_o = do[br]
        grads[t1 + "bias"] = d_o.sum(dim=(0, 2)).view(1, -1)
        grads[t1 + "weight"] = torch.einsum("bwn,bcn->wc", d_o, it["h3"]).unsqueeze(0)
        dh3 = torch.einsum("wc,bwn->bcn", W2, d_o)
        da = dh3 * (it["a"] > 0)
        dt = da.sum(2)                                       # (B,F)
        ds = (da * it["h2n"]).sum(2)
        dh2n = da * it["s"].unsqueeze(2)
        istd_b = 1.0 / torch.sqrt(it["var_b"] + BN_EPS)
        if training:
            dh2pre = _bn_train_backward(dh2n, it["h2n"], istd_b, (0, 2))
        else:
            dh2pre = dh2n * istd_b.view(1, -1, 1)
        grads[t0 + "sd1.weight"] = torch.einsum("bcn,bjn->cj", dh2pre, it["h1"]).unsqueeze(0)
        dh1 = torch.einsum("cj,bcn->bjn", W1, dh2pre)
        dz = dh1 * (it["z"] > 0)
        istd_a = 1.0 / torch.sqrt(it["var_a"] + BN_EPS)
        xhat = (it["h0"] - it["mean_a"].view(1, -1, 1)) * istd_a.view(1, -1, 1)
        gam = P[t0 + "sd0_bn.weight"]
        grads[t0 + "sd0_bn.weight"] = (dz * xhat).sum(dim=(0, 2))
        grads[t0 + "sd0_bn.bias"] = dz.sum(dim=(0, 2))
        dxhat = dz * gam.view(1, -1, 1)
        if training:
            dh0 = _bn_train_backward(dxhat, xhat, istd_a, (0, 2))
        else:
            dh0 = dxhat * istd_a.view(1, -1, 1)
        grads[t0 + "sd0.weight"] = torch.einsum("bcn,bjn->cj", dh0, xk).unsqueeze(0)
        dp[:, keep, :] += torch.einsum("cj,bcn->bjn", W0, dh0)
        # FiLM nets: s = eps + exp(wraw) -> dwraw = ds * exp(wraw) = ds * (s - eps)
        dg = dg + film_net_backward(P, br, "w", g, ds * (it["s"] - eps), training, grads)
        dg = dg + film_net_backward(P, br, "b", g, dt, training, grads)
    return dp, dg, grads


# ------------------------------------------------------------------------------------------------
# The SAME backward, restated in the two-pass + deferred-correction form the CUDA kernels use
# (DESIGN.md "training backward"): BN_a's batch-statistics terms are not applied per point; they
# collapse into per-layer sums (dbeta, E) plus an affine correction  dx[keep] -= cvec + Q x[keep]
# that the next backward step applies when it loads its dy.  Used by tests to validate the algebra.
# ------------------------------------------------------------------------------------------------
def coupling_backward_two_pass(P, p, g, mode, warp_inds, dy_stored, pending=None, dmu_ext=None,
                               dlv_ext=None, training=True):
    """Returns (dp_stored, new_pending, dg, grads).  `pending` = (keep_idx, cvec (k,), Q (k,k)) of the
    layer whose input is this layer's output; dp_true[keep] = dp_stored[keep] - cvec - Q p[keep]."""
    eps = P["eps"]
    p_out, mu, logvar, I = coupling_forward(P, p, g, mode, warp_inds, training, None, return_inter=True)
    keep = I["keep"]
    B, _, N = p.shape
    M = B * N
    dy = dy_stored.clone()
    if pending is not None:
        kidx, cvec, Q = pending
        dy[:, kidx, :] -= cvec.view(1, -1, 1) + torch.einsum("jk,bkn->bjn", Q, p_out[:, kidx, :])
    sigma = I["sigma"]
    if mode == "inverse":
        dp = dy / sigma
        dmu = -dy / sigma
        dl = -dy * p_out * torch.exp(logvar) / (2 * sigma * sigma)
    else:
        dp = dy * sigma
        dmu = dy.clone()
        dl = dy * p * torch.exp(logvar) / (2 * sigma)
    if dmu_ext is not None:
        dmu = dmu + dmu_ext
    if dlv_ext is not None:
        dl = dl + dlv_ext
    do = {"mu": dmu[:, warp_inds, :], "logvar": dl[:, warp_inds, :] / (1 + I["o_logvar"].abs()) ** 2}
    xk = p[:, keep, :].double()
    S1 = xk.sum(dim=(0, 2))                                   # input moments of this layer
    S2 = torch.einsum("bjn,bkn->jk", xk, xk)
    grads = {}
    dg = torch.zeros_like(g)
    k = len(keep)
    cv_new = torch.zeros(k, dtype=torch.float64)
    Q_new = torch.zeros((k, k), dtype=torch.float64)
    for br in BRANCHES:
        it = I[br]
        t0 = "T_%s_0.%s_" % (br, br)
        t1 = "T_%s_1.%s_sd2." % (br, br)
        W0, W1, W2 = P[t0 + "sd0.weight"][0], P[t0 + "sd1.weight"][0], P[t1 + "weight"][0]
        gam = P[t0 + "sd0_bn.weight"]
        d_o = do[br]
        # ---- pass 1: per-(b,c) FiLM sums, W2 grads ----
        dh3 = torch.einsum("wc,bwn->bcn", W2, d_o)
        da = dh3 * (it["a"] > 0)
        dt = da.sum(2)
        ds = (da * it["h2n"]).sum(2)
        grads[t1 + "bias"] = d_o.sum(dim=(0, 2)).view(1, -1)
        grads[t1 + "weight"] = torch.einsum("bwn,bcn->wc", d_o, it["h3"]).unsqueeze(0)
        istd_b = 1.0 / torch.sqrt(it["var_b"] + BN_EPS)
        if training:
            m1 = (it["s"] * dt).sum(0) / M
            m2 = (it["s"] * ds).sum(0) / M
        else:
            m1 = m2 = torch.zeros_like(istd_b)
        # ---- pass 2 ----
        dh2pre = istd_b.view(1, -1, 1) * (da * it["s"].unsqueeze(2) - m1.view(1, -1, 1) - it["h2n"] * m2.view(1, -1, 1))
        grads[t0 + "sd1.weight"] = torch.einsum("bcn,bjn->cj", dh2pre, it["h1"]).unsqueeze(0)
        dh1 = torch.einsum("cj,bcn->bjn", W1, dh2pre)
        dz = dh1 * (it["z"] > 0)
        istd_a = (1.0 / torch.sqrt(it["var_a"] + BN_EPS))
        A0 = (gam * istd_a).view(-1, 1) * W0                   # folded BN_a slope, (F,k)
        T1 = torch.einsum("cj,bcn->bjn", A0, dz)               # per-point term
        dp[:, keep, :] += T1
        dbeta = dz.double().sum(dim=(0, 2))
        E = torch.einsum("bcn,bjn->cj", dz.double(), xk)       # (F,k)
        # ---- finalize (per layer, tiny) ----
        W0d, ia, mean_a, gd = W0.double(), istd_a.double(), it["mean_a"].double(), gam.double()
        dgamma = ia * ((W0d * E).sum(1) - mean_a * dbeta)
        grads[t0 + "sd0_bn.bias"] = dbeta.float()
        grads[t0 + "sd0_bn.weight"] = dgamma.float()
        if training:
            n1 = gd * dbeta / M
            n2 = gd * dgamma / M
        else:
            n1 = n2 = torch.zeros_like(dbeta)
        sx = ia.view(-1, 1) * (W0d @ S2 - mean_a.view(-1, 1) * S1.view(1, -1))      # sum_p xhat_c x_k
        grads[t0 + "sd0.weight"] = (ia.view(-1, 1) * (gd.view(-1, 1) * E - n1.view(-1, 1) * S1.view(1, -1)
                                                      - n2.view(-1, 1) * sx)).float().unsqueeze(0)
        u = (W0d * (ia * n1).view(-1, 1)).sum(0)
        r = (W0d * (ia * ia * n2 * mean_a).view(-1, 1)).sum(0)
        cv_new += u - r
        Q_new += torch.einsum("c,cj,ck->jk", ia * ia * n2, W0d, W0d)
        dg = dg + film_net_backward(P, br, "w", g, ds * (it["s"] - eps), training, grads)
        dg = dg + film_net_backward(P, br, "b", g, dt, training, grads)
    return dp, (keep, cv_new.float(), Q_new.float()), dg, grads


def apply_pending(dp_stored, p, pending):
    """Resolves the deferred BN_a correction on a stored input gradient."""
    kidx, cvec, Q = pending
    dp = dp_stored.clone()
    dp[:, kidx, :] -= cvec.view(1, -1, 1) + torch.einsum("jk,bkn->bjn", Q, p[:, kidx, :])
    return dp


def init_decoder_layers(n_flows, G, F=64, weight_std=0.01, seed=0, device="cpu", requires_grad=True):
    """Random-init parameters of a LocalCondRNVPDecoder with the reference's initialisation
    (flows.py:25-93, layers.py:29-39, decoders.py:49-52): kaiming-uniform SharedDots and nn.Linear defaults,
    N(0, weight_std) final layers with zero bias, BN weight 1 / bias 0 / running stats (0, 1).
    -> (layers = [(param dict keyed like one CondRealNVPFlow3D, warp_inds)], leaves = list of trainable tensors).
    Used by bench.py's reference arm, which must not touch product code."""
    gen = torch.Generator().manual_seed(seed)
    layers, leaves = [], []

    def uni(shape, bound):
        return (torch.rand(shape, generator=gen) * 2.0 - 1.0) * bound

    def leaf(t):
        t = t.to(device)
        if requires_grad:
            t.requires_grad_(True)
            leaves.append(t)
        return t

    for _, warp in decoder_layer_names(n_flows):
        k, w = 3 - len(warp), len(warp)
        P = {"eps": torch.tensor([1e-6], device=device)}
        for br in BRANCHES:
            t0 = "T_%s_0.%s_" % (br, br)
            P[t0 + "sd0.weight"] = leaf(uni((1, F, k), math.sqrt(6.0 / (F * k))))
            P[t0 + "sd0_bn.weight"] = leaf(torch.ones(F))
            P[t0 + "sd0_bn.bias"] = leaf(torch.zeros(F))
            P[t0 + "sd1.weight"] = leaf(uni((1, F, F), math.sqrt(6.0 / (F * F))))
            for bn in (t0 + "sd0_bn", t0 + "sd1_bn"):
                P[bn + ".running_mean"] = torch.zeros(F, device=device)
                P[bn + ".running_var"] = torch.ones(F, device=device)
            for kind in ("w", "b"):
                f = "T_%s_0_cond_%s.%s_sd1_film_%s" % (br, kind, br, kind)
                P[f + "0.weight"] = leaf(uni((F, G), 1.0 / math.sqrt(G)))
                P[f + "0_bn.weight"] = leaf(torch.ones(F))
                P[f + "0_bn.bias"] = leaf(torch.zeros(F))
                P[f + "0_bn.running_mean"] = torch.zeros(F, device=device)
                P[f + "0_bn.running_var"] = torch.ones(F, device=device)
                P[f + "1.weight"] = leaf(torch.randn((F, F), generator=gen) * weight_std)
                P[f + "1.bias"] = leaf(torch.zeros(F))
            t1 = "T_%s_1.%s_sd2." % (br, br)
            P[t1 + "weight"] = leaf(torch.randn((1, w, F), generator=gen) * weight_std)
            P[t1 + "bias"] = leaf(torch.zeros((1, w)))
        layers.append((P, warp))
    return layers, leaves
