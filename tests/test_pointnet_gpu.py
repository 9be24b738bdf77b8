"""Fused eval-mode PointNet encoder + max-pool (dpf_pointnet_eval_forward) against the module's own
torch path = the reference's layer sequence (lib/networks/encoders.py:9-28, models.py:130-131)."""
import os
import sys
import time

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pytestmark = pytest.mark.gpu


def rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def make_encoder(cuda, seed):
    from dpf_nets_b200.lib.networks.encoders import PointNetCloudEncoder
    torch.manual_seed(seed)
    enc = PointNetCloudEncoder(3, 64, [128, 256, 512]).to(cuda)
    gen = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():   # non-trivial BatchNorm state: negative scales, shifted / scaled running statistics
        for m in enc.features:
            if isinstance(m, torch.nn.BatchNorm1d):
                n = m.num_features
                m.weight.copy_((torch.rand(n, generator=gen) * 1.5 + 0.25) * torch.where(torch.rand(n, generator=gen) < 0.2, -1.0, 1.0))
                m.bias.copy_(torch.randn(n, generator=gen) * 0.3)
                m.running_mean.copy_(torch.randn(n, generator=gen) * 0.2)
                m.running_var.copy_(torch.rand(n, generator=gen) * 0.5 + 0.05)
    enc.eval()
    return enc


@pytest.mark.parametrize("B,N", [(32, 2048), (3, 1000), (5, 77), (1, 128), (200, 300), (2, 2500)])
def test_fused_eval_encoder_matches_torch_path(native_lib, cuda, B, N):
    enc = make_encoder(cuda, 3)
    gen = torch.Generator().manual_seed(B * 1000 + N)
    x = (torch.rand((B, 3, N), generator=gen) - 0.5).to(cuda)
    with torch.no_grad():
        got = enc.global_features(x)
        enc.precision = "fp32"
        want = enc.global_features(x)          # torch path (library GEMMs, fp32)
        assert torch.equal(want, torch.max(enc(x), dim=2)[0])
    assert got.shape == (B, 512) and torch.isfinite(got).all()
    err = rel(got, want)
    print("pointnet eval", B, N, "rel err", err)
    assert err < 2e-2          # bf16 tolerance of BASELINE.json north_star; measured ~3e-3
    assert (got >= 0).all()


def test_fused_eval_encoder_is_used_and_fast(native_lib, cuda):
    import ctypes
    enc = make_encoder(cuda, 5)
    x = (torch.rand((32, 3, 2048)) - 0.5).to(cuda)
    n0, n1 = ctypes.c_longlong(0), ctypes.c_longlong(0)
    native_lib.dpf_launch_count(ctypes.byref(n0))
    with torch.no_grad():
        enc.global_features(x)
    native_lib.dpf_launch_count(ctypes.byref(n1))
    assert n1.value - n0.value == 2       # pack + fused kernel
    torch.cuda.synchronize()
    with torch.no_grad():
        t0 = time.perf_counter()
        for _ in range(20):
            enc.global_features(x)
        torch.cuda.synchronize()
        fused = (time.perf_counter() - t0) / 20
        enc.precision = "fp32"
        enc.global_features(x)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            enc.global_features(x)
        torch.cuda.synchronize()
        lib_path = (time.perf_counter() - t0) / 5
    print("pointnet eval 32x2048: fused %.3f ms, library path %.3f ms" % (fused * 1e3, lib_path * 1e3))
    assert fused < lib_path


def test_eval_mode_autograd_and_no_grad_train_keep_the_library_path(native_lib, cuda):
    enc = make_encoder(cuda, 7)
    x = (torch.rand((4, 3, 256)) - 0.5).to(cuda)
    out = enc.global_features(x)            # eval mode with grad enabled -> torch path, differentiable
    assert out.requires_grad
    enc.train()
    with torch.no_grad():
        assert torch.equal(enc.global_features(x), torch.max(enc(x), dim=2)[0])


@pytest.mark.parametrize("B,N", [(1, 1), (2, 64), (3, 77), (4, 1000), (5, 2049), (32, 2048)])
def test_pool_statistics_kernel_vs_torch(native_lib, cuda, B, N):
    """dpf_pointnet_pool_forward (W h2 on the tensor cores with split bf16 operands, one channel per TMEM lane): batch sums,
    max / min over the points and their indices against the fp32 torch definition.  Values 2e-5 relative to the channel
    scale (bf16 hi + lo operands: ~2^-17 per operand); an index may differ only where the two candidates are that close."""
    from dpf_nets_b200.ops.pointnet_pool import _pool_stats
    g = torch.Generator().manual_seed(B * 1000 + N)
    h2 = torch.relu(torch.randn((B, 256, N), generator=g)).to(cuda)
    W = (torch.randn((512, 256), generator=g) * 0.08).to(cuda)
    mean, var, vmax, vmin, imax, imin = _pool_stats(h2, W)
    h = torch.matmul(W.double(), h2.double())                     # (B,512,N) truth
    scale = h.abs().max().clamp_min(1e-6)                         # one scale for all channels (split-bf16 products: ~1e-5 of it)
    assert ((mean - h.mean((0, 2))).abs() <= 1e-4 * scale).all()
    tvar = h.var((0, 2), unbiased=False)
    assert ((var - tvar).abs() <= 2e-4 * tvar + 1e-4 * scale * scale * 1e-3).all()
    tmax, tmin = h.max(2)[0], h.min(2)[0]
    assert ((vmax.double() - tmax).abs() <= 1e-4 * scale).all() and ((vmin.double() - tmin).abs() <= 1e-4 * scale).all()
    # the selected point attains the extremum (up to the same tolerance)
    assert ((torch.gather(h, 2, imax.unsqueeze(2)).squeeze(2) - tmax).abs() <= 2e-4 * scale).all()
    assert ((torch.gather(h, 2, imin.unsqueeze(2)).squeeze(2) - tmin).abs() <= 2e-4 * scale).all()
    assert int(imax.min()) >= 0 and int(imax.max()) < N and int(imin.min()) >= 0 and int(imin.max()) < N


@pytest.mark.parametrize("fused_layers", [False, True])
@pytest.mark.parametrize("B,N", [(4, 300), (32, 2048)])
def test_train_mode_fused_last_layer_vs_library_path(native_lib, cuda, B, N, fused_layers, monkeypatch):
    """Train-mode global_features: fused last layer + BatchNorm (batch statistics) + ReLU + max-pool with the analytic
    backward (ops/pointnet_pool.py) against the library path of the same module (SharedDot -> BatchNorm1d -> ReLU -> max,
    the reference's chain: encoders.py:9-28, models.py:130-131; pinned to the reference on CPU): pooled features,
    running statistics and the gradients of EVERY encoder parameter."""
    from dpf_nets_b200.lib.networks.encoders import PointNetCloudEncoder
    # fused_layers: the whole encoder as one autograd function over the library's kernels (ops/pointnet_train.py: narrow layers
    # forward + backward as tcgen05 kernels); False: only the last layer + max-pool fused (ops/pointnet_pool.py)
    monkeypatch.setattr(PointNetCloudEncoder, "fused_layers", fused_layers)
    enc = make_encoder(cuda, 11)
    with torch.no_grad():                     # both signs of gamma in the last BatchNorm: max- and min-selected channels
        enc.features.sd2_bn.weight.mul_(torch.where(torch.rand(512, device=cuda) < 0.3, -1.0, 1.0))
    x = (torch.rand((B, 3, N), generator=torch.Generator().manual_seed(5)) - 0.5).to(cuda)
    cot = torch.randn((B, 512), generator=torch.Generator().manual_seed(6)).to(cuda)
    sd0 = {k: v.clone() for k, v in enc.state_dict().items()}
    res = {}
    for prec in ("fp32", "auto"):
        enc.load_state_dict(sd0)
        enc.train()
        enc.precision = prec
        enc.zero_grad()
        out = enc.global_features(x)
        (out * cot).sum().backward()
        res[prec] = (out.detach().clone(), {k: p.grad.clone() for k, p in enc.named_parameters()},
                     {k: v.clone() for k, v in enc.state_dict().items() if "running" in k or "num_batches" in k})
    lib, fused = res["fp32"], res["auto"]
    assert rel(fused[0], lib[0]) < 1e-4
    for k, v in lib[2].items():
        assert (torch.equal(fused[2][k], v) if "num_batches" in k else rel(fused[2][k], v) < 1e-4), k
    # Gradients: the max-pool routes each (shape, channel) cotangent to ONE point; where two points of a cloud are within
    # fp32 rounding of each other (a few hundred of the 16 384 (shape, channel) pairs at 2048 points) the kernel's split-bf16
    # products and cuBLAS' fp32 products may pick different ones - both are argmaxes to fp32 accuracy and the gradient
    # legitimately differs there.  Gate: L2-relative error (a few flipped selections barely move it); the exact algebra is
    # checked in float64 at full width below and on the CPU (tests/test_models_host.py).
    def l2(a, b):
        return float((a - b).norm() / b.norm().clamp_min(1e-30))
    worst = max((l2(fused[1][k], v), k) for k, v in lib[1].items())
    assert worst[0] < (2e-3 if N <= 300 else 3e-2), worst


def test_pooled_last_layer_float64_on_gpu_at_full_width(native_lib, cuda, monkeypatch):
    """The op's forward selection logic and analytic backward at 2048 points per cloud in float64 (statistics pass replaced
    by its torch definition, exact argmax) against autograd of the module chain in float64: 1e-9."""
    from dpf_nets_b200.ops import pointnet_pool as pp

    def stats(h2, W):
        h = torch.matmul(W, h2)
        vmax, imax = h.max(2)
        vmin, imin = h.min(2)
        return h.mean((0, 2)), h.var((0, 2), unbiased=False), vmax, vmin, imax, imin
    monkeypatch.setattr(pp, "_pool_stats", stats)
    g = torch.Generator().manual_seed(2)
    B, N = 6, 2048
    h2 = torch.relu(torch.randn((B, 256, N), generator=g, dtype=torch.float64)).to(cuda).requires_grad_(True)
    W = (torch.randn((512, 256), generator=g, dtype=torch.float64) * 0.08).to(cuda).requires_grad_(True)
    bn = torch.nn.BatchNorm1d(512).double().to(cuda)
    with torch.no_grad():
        bn.weight.copy_(torch.randn(512, generator=g, dtype=torch.float64))
        bn.bias.copy_(0.3 * torch.randn(512, generator=g, dtype=torch.float64))
    bn2 = torch.nn.BatchNorm1d(512).double().to(cuda)
    bn2.load_state_dict(bn.state_dict())
    cot = torch.randn((B, 512), generator=g, dtype=torch.float64).to(cuda)
    ref = torch.max(torch.relu(bn(torch.matmul(W, h2))), dim=2)[0]
    gr = torch.autograd.grad((ref * cot).sum(), [h2, W, bn.weight, bn.bias])
    out = pp.pooled_bn_relu_max(h2, W, bn2)
    go = torch.autograd.grad((out * cot).sum(), [h2, W, bn2.weight, bn2.bias])
    assert rel(out, ref) < 1e-12
    for a, b in zip(go, gr):
        assert rel(a, b) < 1e-9


def test_train_mode_fused_last_layer_is_faster(native_lib, cuda):
    enc = make_encoder(cuda, 3).train()
    x = (torch.rand((32, 3, 2048)) - 0.5).to(cuda)
    cot = torch.randn((32, 512), device=cuda)

    def step():
        enc.zero_grad()
        (enc.global_features(x) * cot).sum().backward()
    times = {}
    for prec in ("fp32", "auto"):
        enc.precision = prec
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10):
            step()
        b.record()
        torch.cuda.synchronize()
        times[prec] = a.elapsed_time(b) / 10
    print("pointnet train fwd+bwd 32x2048: library path %.3f ms, fused last layer %.3f ms" % (times["fp32"], times["auto"]))
    assert times["auto"] < times["fp32"]


def _tables(gen, C, cuda):
    sc = (torch.rand(C, generator=gen) + 0.5) * torch.where(torch.rand(C, generator=gen) < 0.2, -1.0, 1.0)
    sh = torch.randn(C, generator=gen) * 0.3
    return sc.to(cuda), sh.to(cuda)


@pytest.mark.parametrize("B,N", [(3, 200), (2, 333), (5, 1024)])
def test_pointnet_layer_kernels_vs_float64(native_lib, cuda, B, N):
    """csrc/pointnet_layers.cu unit by unit against float64 torch on the same tables: the three operand loaders (layer 0 from
    the coordinates, BatchNorm + ReLU on load, BatchNorm + ReLU backward on load), the GEMM kernel (outputs, per-work-item
    statistics merged like the product does, row offsets, padded output rows), the weight-gradient / Gram kernel and the two
    streaming reductions.  Ragged tiles (N % 64 != 0) and the unaligned scalar path (N % 4 != 0) included."""
    from dpf_nets_b200 import _lib
    from dpf_nets_b200.ops import pointnet_train as pt
    gen = torch.Generator().manual_seed(B * 1000 + N)
    r = lambda *s: torch.randn(s, generator=gen)
    x = (torch.rand((B, 3, N), generator=gen) - 0.5).to(cuda)
    tol = 2e-5

    def relerr(a, b):
        return float((a.double() - b).abs().max() / b.abs().max().clamp_min(1e-30))

    # ---- loader 0 (layer 0 from x), K = 64 -> 128, statistics ----
    a = (r(64, 3) * 1.5).to(cuda)
    c = (r(64) * 0.3).to(cuda)
    tab0 = pt._table(a[:, 0], a[:, 1], a[:, 2], c)
    W1 = (r(128, 64) * 0.2).to(cuda)
    Z1, st1 = pt._gemm(pt.LD_X3, 64, x, None, tab0, pt._image(W1, 128, 64, False), B, N, 128, want_stat=True)
    A0 = torch.relu(torch.einsum('ci,bin->bcn', a.double(), x.double()) + c.double().view(1, -1, 1))
    Z1_ref = torch.einsum('oc,bcn->bon', W1.double(), A0)
    assert relerr(Z1, Z1_ref) < tol
    mean1, var1 = pt._merge_stats(st1)
    assert relerr(mean1, Z1_ref.mean((0, 2))) < 1e-5 and relerr(var1, Z1_ref.var((0, 2), unbiased=False)) < 1e-5
    # ---- loader 1 (BatchNorm + ReLU on load), K = 128 -> 256, statistics ----
    sc1, sh1 = _tables(gen, 128, cuda)
    mu1 = Z1_ref.mean((0, 2))
    tab1 = pt._table(sc1, sh1, None, None, None, mu1)
    W2 = (r(256, 128) * 0.15).to(cuda)
    Z1f = Z1_ref.float().contiguous()
    Z2, st2 = pt._gemm(pt.LD_AFFINE, 128, Z1f, None, tab1, pt._image(W2, 256, 128, False), B, N, 256, want_stat=True)
    A1 = torch.relu(Z1f.double() * sc1.double().view(1, -1, 1) + sh1.double().view(1, -1, 1))
    Z2_ref = torch.einsum('oc,bcn->bon', W2.double(), A1)
    assert relerr(Z2, Z2_ref) < tol
    mean2, var2 = pt._merge_stats(st2)
    assert relerr(mean2, Z2_ref.mean((0, 2))) < 1e-5 and relerr(var2, Z2_ref.var((0, 2), unbiased=False)) < 1e-5
    # ---- loader 2 (BatchNorm + ReLU backward on load): dgrads K = 256 -> 128 and K = 128 -> 64 (padded rows), row offsets ----
    Z2f = Z2_ref.float().contiguous()
    sc2, sh2 = _tables(gen, 256, cuda)
    mu2 = Z2_ref.mean((0, 2))
    dA2 = r(B, 256, N).to(cuda)
    g2, gm12, k22 = (r(256) * 0.5).to(cuda), (r(256) * 0.1).to(cuda), (r(256) * 0.1).to(cuda)
    tabb2 = pt._table(sc2, sh2, g2, gm12, k22, mu2)

    def bnbwd(dA, Z, sc, sh, g, gm1, k2, mu):
        v = lambda t: t.double().view(1, -1, 1)
        y = Z.double() * v(sc) + v(sh)
        return torch.where(y > 0, v(g) * dA.double(), torch.zeros_like(y)) - v(gm1) - v(k2) * (Z.double() - v(mu))
    dZ2 = bnbwd(dA2, Z2f, sc2, sh2, g2, gm12, k22, mu2)
    roff = (r(128) * 0.2).to(cuda)
    dA1, _ = pt._gemm(pt.LD_BNBWD, 256, dA2, Z2f, tabb2, pt._image(W2, 128, 256, True), B, N, 128, row_off=roff)
    dA1_ref = torch.einsum('oc,bon->bcn', W2.double(), dZ2) + roff.double().view(1, -1, 1)
    assert relerr(dA1, dA1_ref) < tol
    g1, gm11, k21 = (r(128) * 0.5).to(cuda), (r(128) * 0.1).to(cuda), (r(128) * 0.1).to(cuda)
    tabb1 = pt._table(sc1, sh1, g1, gm11, k21, mu1)
    dA1f = dA1_ref.float().contiguous()
    dZ1 = bnbwd(dA1f, Z1f, sc1, sh1, g1, gm11, k21, mu1)
    dA0, _ = pt._gemm(pt.LD_BNBWD, 128, dA1f, Z1f, tabb1, pt._image(W1, 64, 128, True), B, N, 64)
    assert relerr(dA0, torch.einsum('oc,bon->bcn', W1.double(), dZ1)) < tol
    # ---- weight gradients and the Gram matrix ----
    dW2 = pt._wgrad(256, 128, 0, pt.LD_AFFINE, dA2, Z2f, tabb2, Z1f, tab1, B, N)
    assert relerr(dW2, torch.einsum('bon,bcn->oc', dZ2, A1)) < tol
    dW1 = pt._wgrad(128, 64, 0, pt.LD_X3, dA1f, Z1f, tabb1, x, tab0, B, N)
    assert relerr(dW1, torch.einsum('bon,bcn->oc', dZ1, A0)) < tol
    sub = (r(256) * 0.2).to(cuda)
    tabg = pt._table(sc2, sh2, sub)
    G = pt._wgrad(256, 256, 1, pt.LD_AFFINE, Z2f, None, tabg, None, None, B, N)
    hc = torch.relu(Z2f.double() * sc2.double().view(1, -1, 1) + sh2.double().view(1, -1, 1)) - sub.double().view(1, -1, 1)
    assert relerr(G, torch.einsum('bin,bjn->ij', hc, hc)) < tol
    # ---- streaming reductions ----
    s2 = torch.zeros((256, 2), dtype=torch.float64, device=cuda)
    _lib.call("dpf_pointnet_bn_bwd_sums", dA2, Z2f, tabb2, B, 256, N, s2, device=cuda)
    y2 = Z2f.double() * sc2.double().view(1, -1, 1) + sh2.double().view(1, -1, 1)
    dm = torch.where(y2 > 0, dA2.double(), torch.zeros_like(y2))
    assert relerr(s2[:, 0], dm.sum((0, 2))) < 1e-5 and relerr(s2[:, 1], (dm * (Z2f.double() - mu2.view(1, -1, 1))).sum((0, 2))) < 1e-5
    s0 = torch.zeros((64, 4), dtype=torch.float64, device=cuda)
    dA0f = dA0.contiguous()
    _lib.call("dpf_pointnet_layer0_bwd_sums", dA0f, x, tab0, B, 64, N, s0, device=cuda)
    y0 = torch.einsum('ci,bin->bcn', a.double(), x.double()) + c.double().view(1, -1, 1)
    dm0 = torch.where(y0 > 0, dA0f.double(), torch.zeros_like(y0))
    assert relerr(s0[:, 0], dm0.sum((0, 2))) < 1e-5
    assert relerr(s0[:, 1:], torch.einsum('bcn,bjn->cj', dm0, x.double())) < 1e-5


def test_pointnet_train_function_float64_truth(native_lib, cuda):
    """The fused train-mode encoder (ops/pointnet_train.py) and the library chain of the same module, both against the SAME
    module evaluated in float64: the fused path may be no further from the truth than a small multiple of the library's own
    fp32 result in the pooled features, and within the split-bf16 gradient gate (2e-2 in L2, see below) for all twelve
    parameter tensors."""
    import copy
    B, N = 8, 1024
    enc = make_encoder(cuda, 21)
    x = (torch.rand((B, 3, N), generator=torch.Generator().manual_seed(8)) - 0.5).to(cuda)
    cot = torch.randn((B, 512), generator=torch.Generator().manual_seed(9)).to(cuda)
    sd0 = {k: v.clone() for k, v in enc.state_dict().items()}

    def run(module, xin, cotin, prec):
        module.train()
        module.precision = prec
        module.zero_grad()
        out = module.global_features(xin)
        (out * cotin).sum().backward()
        return out.detach(), {k: p.grad.clone() for k, p in module.named_parameters()}
    enc64 = copy.deepcopy(enc).double()
    truth = run(enc64, x.double(), cot.double(), "fp32")
    res = {}
    for prec in ("fp32", "auto"):
        enc.load_state_dict(sd0)
        res[prec] = run(enc, x, cot, prec)

    def l2(a, b):
        return float((a.double() - b).norm() / b.norm().clamp_min(1e-30))
    errs = {k: (l2(res["fp32"][1][k], truth[1][k]), l2(res["auto"][1][k], truth[1][k])) for k in truth[1]}
    print("gradient L2 errors vs float64 (library fp32, fused):", {k: ("%.1e" % a, "%.1e" % b) for k, (a, b) in errs.items()})
    # What bounds the fused path: its forward is accurate to the split-bf16 GEMMs' ~1e-5, so a fraction ~1e-5 of the ReLU masks
    # (activations within that distance of 0) and a few max-pool selections flip relative to float64; each flip changes a
    # gradient term by O(1), i.e. an L2-relative gradient error ~sqrt(1e-5) = 3e-3 in every layer below the flip (measured
    # 2e-3 .. 6e-3; the last layer's own gradients, which see no mask, are at 1e-5).  Same gate as the decoder's bf16x3 gradients.
    for k, (e_lib, e_fused) in errs.items():
        assert e_fused < max(2e-2, 4 * e_lib), (k, e_fused, e_lib)
    assert l2(res["auto"][0], truth[0]) < max(1e-4, 4 * l2(res["fp32"][0], truth[0]))
