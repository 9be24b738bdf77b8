"""GPU: the fused latent-side blocks (csrc/latent.cu: BatchNorm1d + Swish, RealNVPFlow transform) against the plain module
chains of the same classes (torch ops; pinned to the reference on the CPU by tests/test_models_host.py with goldens generated
from lib/networks/flows.py:163-243, decoders.py:7-38, encoders.py:31-83): outputs, running statistics and every gradient,
train and eval mode, both flow directions, both coupling patterns, G = 16 / 128 / 512."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def rel(a, b):
    return float((a.detach() - b.detach()).abs().max() / b.detach().abs().max().clamp_min(1e-30))


def _run(module, fused_classes, fused, fn):
    old = [c.fused for c in fused_classes]
    for c in fused_classes:
        c.fused = fused
    try:
        return fn(module)
    finally:
        for c, o in zip(fused_classes, old):
            c.fused = o


@pytest.mark.parametrize("fused_layer", [True, False])
@pytest.mark.parametrize("G,B", [(16, 5), (128, 32), (512, 32), (128, 64), (128, 33), (128, 2)])
@pytest.mark.parametrize("mode", ["inverse", "direct"])
@pytest.mark.parametrize("training", [True, False])
def test_global_rnvp_decoder_fused_vs_module_chain(native_lib, cuda, G, B, mode, training, fused_layer, monkeypatch):
    from dpf_nets_b200.lib.networks.decoders import GlobalRNVPDecoder
    from dpf_nets_b200.lib.networks.flows import RealNVPFlow
    # fused_layer: every coupling layer as ONE kernel forward and ONE backward (csrc/latent_flow.cu; B <= 64, kept / warped
    # widths multiples of 32 - G = 16 falls back to the block kernels); False: BatchNorm + Swish and the transform as
    # kernels, the four Linear layers through the library
    monkeypatch.setattr(RealNVPFlow, "fused_layer", fused_layer)
    torch.manual_seed(G + B)
    m = GlobalRNVPDecoder(3, 64 if G < 128 else 128, G, weight_std=0.05).to(cuda)
    with torch.no_grad():
        for n, p in m.named_parameters():
            if "bn.weight" in n:
                p.add_(0.2 * torch.randn_like(p))
            if "bn.bias" in n:
                p.add_(0.2 * torch.randn_like(p))
    sd0 = {k: v.clone() for k, v in m.state_dict().items()}
    g0 = torch.randn((B, G), generator=torch.Generator().manual_seed(1)).to(cuda)
    cots = [torch.randn((6, B, G), generator=torch.Generator().manual_seed(2 + i)).to(cuda) for i in range(3)]

    def fn(mod):
        mod.load_state_dict(sd0)
        mod.train(training)
        mod.zero_grad()
        g = g0.clone().requires_grad_(True)
        gs, mus, lvs = mod(g, mode=mode)
        loss = sum((torch.stack(t) * c).sum() for t, c in zip((gs, mus, lvs), cots))
        loss.backward()
        return (torch.stack(gs), torch.stack(mus), torch.stack(lvs), g.grad.clone(),
                {k: p.grad.clone() for k, p in mod.named_parameters()},
                {k: v.clone() for k, v in mod.state_dict().items() if "running" in k or "num_batches" in k})
    ref = _run(m, [RealNVPFlow], False, fn)
    got = _run(m, [RealNVPFlow], True, fn)
    # two-row batches: batch-statistics BatchNorm divides by |x0 - x1|, fp32 rounding on both sides shows up at ~1e-3
    tol_o, tol_p = (2e-4, 2e-3) if B >= 5 else (3e-3, 2e-2)
    for a, b, what in zip(got[:4], ref[:4], ("g", "mu", "logvar", "dg")):
        assert rel(a, b) < tol_o, (what, rel(a, b))
    worst = max((rel(got[4][k], v), k) for k, v in ref[4].items())
    assert worst[0] < tol_p, worst
    for k, v in ref[5].items():
        assert (torch.equal(got[5][k], v) if "num_batches" in k else rel(got[5][k], v) < 1e-5), k


@pytest.mark.parametrize("deterministic", [False, True])
@pytest.mark.parametrize("training", [True, False])
def test_feature_encoder_fused_vs_module_chain(native_lib, cuda, deterministic, training):
    from dpf_nets_b200.lib.networks.encoders import FeatureEncoder
    torch.manual_seed(3)
    m = FeatureEncoder(2, 512, 128, deterministic=deterministic, mu_weight_std=0.1, logvar_weight_std=0.1).to(cuda)
    sd0 = {k: v.clone() for k, v in m.state_dict().items()}
    x0 = torch.randn((32, 512), generator=torch.Generator().manual_seed(4)).to(cuda)
    cot = torch.randn((2, 32, 128), generator=torch.Generator().manual_seed(5)).to(cuda)

    def fn(mod):
        mod.load_state_dict(sd0)
        mod.train(training)
        mod.zero_grad()
        x = x0.clone().requires_grad_(True)
        out = mod(x)
        outs = (out,) if deterministic else out
        sum((o * c).sum() for o, c in zip(outs, cot)).backward()
        return ([o.detach().clone() for o in outs], x.grad.clone(), {k: p.grad.clone() for k, p in mod.named_parameters()},
                {k: v.clone() for k, v in mod.state_dict().items() if "running" in k or "num_batches" in k})
    ref = _run(m, [FeatureEncoder], False, fn)
    got = _run(m, [FeatureEncoder], True, fn)
    for a, b in zip(got[0], ref[0]):
        assert rel(a, b) < 1e-4
    assert rel(got[1], ref[1]) < 1e-3
    worst = max((rel(got[2][k], v), k) for k, v in ref[2].items())
    assert worst[0] < 1e-3, worst
    for k, v in ref[3].items():
        assert (torch.equal(got[3][k], v) if "num_batches" in k else rel(got[3][k], v) < 1e-5), k


@pytest.mark.parametrize("fused_layer,n_launch", [(True, 1), (False, 3)])
def test_latent_blocks_launch_counts(native_lib, cuda, fused_layer, n_launch, monkeypatch):
    """One coupling layer of the latent flow = ONE kernel forward and ONE backward (csrc/latent_flow.cu); with the block
    kernels only: 2 x bn_swish + 1 transform kernel forward, the same backward."""
    import ctypes
    from dpf_nets_b200.lib.networks.flows import RealNVPFlow
    monkeypatch.setattr(RealNVPFlow, "fused_layer", fused_layer)
    m = RealNVPFlow(128, 128, warp_inds=list(range(0, 128, 2))).to(cuda).train()
    g = torch.randn((32, 128), device=cuda, requires_grad=True)
    n0, n1, n2 = ctypes.c_longlong(0), ctypes.c_longlong(0), ctypes.c_longlong(0)
    native_lib.dpf_launch_count(ctypes.byref(n0))
    out = m(g, mode="inverse")
    native_lib.dpf_launch_count(ctypes.byref(n1))
    (out[0].sum() + out[2].sum()).backward()
    native_lib.dpf_launch_count(ctypes.byref(n2))
    assert n1.value - n0.value == n_launch and n2.value - n1.value == n_launch


@pytest.mark.parametrize("B,F", [(2, 7), (32, 128), (33, 100), (64, 512), (65, 40), (256, 64)])
def test_bn_swish_kernel_shapes(native_lib, cuda, B, F):
    """Every instantiation of the fused BatchNorm1d + Swish pair (4 / 8 / 32 rows per thread, ragged column tiles) against
    torch: values, saved running statistics, gradients."""
    from dpf_nets_b200.ops.latent import bn_swish
    torch.manual_seed(B * 31 + F)
    x0 = torch.randn((B, F), device=cuda)
    cot = torch.randn((B, F), device=cuda)
    bn_a, bn_b = torch.nn.BatchNorm1d(F).to(cuda), torch.nn.BatchNorm1d(F).to(cuda)
    with torch.no_grad():
        bn_a.weight.copy_(1 + 0.3 * torch.randn(F, device=cuda)); bn_a.bias.copy_(0.3 * torch.randn(F, device=cuda))
    bn_b.load_state_dict(bn_a.state_dict())
    for training in (True, False):
        bn_a.train(training); bn_b.train(training)
        xa, xb = x0.clone().requires_grad_(True), x0.clone().requires_grad_(True)
        za = bn_a(xa)
        ya = za * torch.sigmoid(za)
        yb = bn_swish(xb, bn_b)
        (ya * cot).sum().backward()
        (yb * cot).sum().backward()
        # two-row batches have istd ~ 1/|x0-x1|: the fp32 rounding of both sides shows up at a few 1e-4
        assert rel(yb, ya) < 1e-5 and rel(xb.grad, xa.grad) < (1e-4 if B >= 8 else 1e-3)
        assert rel(bn_b.weight.grad, bn_a.weight.grad) < 1e-4 and rel(bn_b.bias.grad, bn_a.bias.grad) < 1e-4
        assert rel(bn_b.running_mean, bn_a.running_mean) < 1e-5 and rel(bn_b.running_var, bn_a.running_var) < 1e-5
        bn_a.zero_grad(); bn_b.zero_grad()
