"""GPU parity of the coupling stack (fp32 path) through the module API -> C ABI, against the
golden vectors produced by the reference (tests/golden) and against the CPU oracle."""
import os

import pytest
import torch

from oracle import flow_oracle as fo

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL32 = 1e-4   # north_star: coupling outputs / per-point log-det within 1e-4 relative in fp32 mode


def rel(a, b):
    a, b = a.detach().cpu(), b.detach().cpu()
    a = a.to(b.dtype) if b.dtype == torch.float64 else a.float()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def load(name):
    return torch.load(os.path.join(GOLD, name), weights_only=False)


@pytest.fixture(scope="module")
def mods(native_lib, cuda):
    from dpf_nets_b200.lib.networks import flows, decoders
    return flows, decoders


@pytest.mark.parametrize("name", ["coupling_w0.pt", "coupling_w02.pt", "coupling_w1_g128.pt"])
def test_coupling_layer_eval_and_train_forward(mods, cuda, name):
    flows, _ = mods
    fx = load(name)
    m = flows.CondRealNVPFlow3D(64, fx["G"], warp_inds=fx["warp"]).to(cuda)
    m.load_state_dict(fx["state"])
    p, g = fx["p"].to(cuda), fx["g"].to(cuda)
    m.eval()
    with torch.no_grad():
        for mode in ("direct", "inverse"):
            out = m(p, g, mode=mode)
            for a, b in zip(out, fx["eval_" + mode]):
                assert rel(a, b) < TOL32, (mode, rel(a, b))
        y = m(p, g, mode="direct")[0]
        x = m(y, g, mode="inverse")[0]
        assert rel(x, p) < 1e-5                      # round trip
    for mode in ("inverse", "direct"):
        m.load_state_dict(fx["state"])
        m.train()
        with torch.no_grad():
            out = m(p, g, mode=mode)
        t = fx["train_" + mode]
        for a, b in zip(out, t["out"]):
            assert rel(a, b) < TOL32, (mode, rel(a, b))
        sd = m.state_dict()
        for k, v in t["state_after"].items():
            if "running" in k:
                assert rel(sd[k], v) < 1e-4, (k, rel(sd[k], v))
            elif "num_batches" in k:
                assert int(sd[k]) == int(v), k
        keep = [c for c in range(3) if c not in fx["warp"]]
        assert (out[1][:, keep] == 0).all() and (out[2][:, keep] == 0).all()


def test_decoder_stack_lists(mods, cuda):
    _, decoders = mods
    fx = load("decoder_f2.pt")
    m = decoders.LocalCondRNVPDecoder(fx["n_flows"], 64, fx["G"]).to(cuda)
    m.load_state_dict(fx["state"])
    p, g = fx["p"].to(cuda), fx["g"].to(cuda)
    m.eval()
    with torch.no_grad():
        for mode in ("direct", "inverse"):
            ps, mus, lvs = m(p, g, mode=mode)
            e = fx["eval_" + mode]
            assert len(ps) == 6 and ps[0].shape == p.shape
            assert rel(torch.stack(list(ps)), e["ps"]) < TOL32
            assert rel(torch.stack(list(mus)), e["mus"]) < TOL32
            assert rel(torch.stack(list(lvs)), e["logvars"]) < TOL32
    m.train()
    with torch.no_grad():
        ps, mus, lvs = m(p, g, mode="inverse")
    t = fx["train_inverse"]
    # Batch-stat BN over B=3 shapes amplifies fp32 rounding through the stack: the reference's own
    # fp32 output is ~1.5e-4 away from the fp64 truth on this fixture, so the gate is "no further
    # from the fp64 truth (oracle in double) than twice the reference itself, or 1e-4".
    names = fo.decoder_layer_names(fx["n_flows"])
    l64 = [({k[len(pre):]: (v.double() if v.is_floating_point() else v) for k, v in fx["state"].items()
             if k.startswith(pre)}, w) for pre, w in names]
    tp, _, tl = fo.decoder_forward(l64, fx["p"].double(), fx["g"].double(), "inverse", training=True)
    tp, tl = torch.stack(tp), torch.stack(tl)
    assert rel(ps.stacked.double(), tp) < max(TOL32, 2 * rel(t["ps"].double(), tp))
    assert rel(lvs.stacked.double(), tl) < max(TOL32, 2 * rel(t["logvars"].double(), tl))
    pc = p.cpu()
    nll = fo.point_flow_nll([x.cpu() for x in ps] + [pc], [torch.zeros_like(pc)] + [x.cpu() for x in mus],
                            [torch.full_like(pc, t["base_logvar"])] + [x.cpu() for x in lvs])
    assert abs(nll.item() - t["nll"].item()) < 5e-3 * abs(t["nll"].item())   # total NLL within 0.5 %
    sd = m.state_dict()
    for k, v in t["state_after"].items():
        if "running" in k:
            assert rel(sd[k], v) < 2e-4, (k, rel(sd[k], v))


def test_decoder_full_depth_vs_oracle(mods, cuda):
    """63 layers, G=128 (airplane/chair config), random non-trivial weights, eval both modes."""
    _, decoders = mods
    torch.manual_seed(0)
    m = decoders.LocalCondRNVPDecoder(21, 64, 128)
    with torch.no_grad():
        v = m.named_views()
        g0 = torch.Generator().manual_seed(5)
        for k, t in v.items():
            if k.endswith("sd2.weight") or k.endswith("film_w1.weight") or k.endswith("film_b1.weight"):
                t.copy_(torch.randn(t.shape, generator=g0) * 0.05)
            if "running_mean" in k:
                t.copy_(torch.randn(t.shape, generator=g0) * 0.1)
            if "running_var" in k:
                t.copy_(torch.rand(t.shape, generator=g0) + 0.5)
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    layers = [({k[len(pre):]: x for k, x in sd.items() if k.startswith(pre)}, w)
              for pre, w in fo.decoder_layer_names(21)]
    gen = torch.Generator().manual_seed(1)
    p = torch.rand((2, 3, 300), generator=gen) - 0.5
    g = torch.randn((2, 128), generator=gen)
    m = m.to(cuda).eval()
    for mode in ("direct", "inverse"):
        ops, omus, olvs = fo.decoder_forward(layers, p, g, mode)
        with torch.no_grad():
            ps, mus, lvs = m(p.to(cuda), g.to(cuda), mode=mode)
        assert rel(ps.stacked, torch.stack(ops)) < TOL32
        assert rel(lvs.stacked, torch.stack(olvs)) < TOL32
        assert rel(mus.stacked, torch.stack(omus)) < TOL32


def test_rejects_cpu_tensors_and_bad_width(mods):
    flows, _ = mods
    from dpf_nets_b200 import DpfNativeError
    m = flows.CondRealNVPFlow3D(64, 8)
    with pytest.raises(DpfNativeError):
        m(torch.zeros(2, 3, 5), torch.zeros(2, 8))
    with pytest.raises(ValueError):
        flows.CondRealNVPFlow3D(32, 8)


# ---------------------------------------------------------------------------------------------
# backward
# ---------------------------------------------------------------------------------------------
TOLG = 1e-3   # gradients: fp32 re-association noise through BN backward (reference vs its own fp64 ~3e-4)


@pytest.mark.parametrize("name", ["coupling_w0.pt", "coupling_w02.pt", "coupling_w1_g128.pt"])
@pytest.mark.parametrize("mode", ["inverse", "direct"])
def test_coupling_layer_backward_vs_reference_autograd(mods, cuda, name, mode):
    flows, _ = mods
    fx = load(name)
    t = fx["train_" + mode]
    m = flows.CondRealNVPFlow3D(64, fx["G"], warp_inds=fx["warp"]).to(cuda)
    m.load_state_dict(fx["state"])
    m.train()
    p = fx["p"].to(cuda).requires_grad_(True)
    g = fx["g"].to(cuda).requires_grad_(True)
    p_out, mu, lv = m(p, g, mode=mode)
    cy, cm, cl = [c.to(cuda) for c in t["cot"]]
    ((p_out * cy).sum() + (mu * cm).sum() + (lv * cl).sum()).backward()
    assert rel(p.grad, t["dp"]) < TOLG, rel(p.grad, t["dp"])
    assert rel(g.grad, t["dg"]) < TOLG, rel(g.grad, t["dg"])
    gv = m.named_views(grad=True)
    worst = max((rel(gv[k], v), k) for k, v in t["grads"].items())
    assert worst[0] < TOLG, worst


def test_decoder_nll_backward_vs_reference_autograd(mods, cuda):
    _, decoders = mods
    from dpf_nets_b200.lib.networks.losses import PointFlowNLL
    fx = load("decoder_f2.pt")
    t = fx["train_inverse"]
    m = decoders.LocalCondRNVPDecoder(fx["n_flows"], 64, fx["G"]).to(cuda)
    m.load_state_dict(fx["state"])
    m.train()
    p = fx["p"].to(cuda)
    g = fx["g"].to(cuda).requires_grad_(True)
    ps, mus, lvs = m(p, g, mode="inverse")
    base_mu, base_lv = torch.zeros_like(p), torch.full_like(p, t["base_logvar"])
    nll = PointFlowNLL()(decoders.prepend(None, ps)[1:] + [p], decoders.prepend(base_mu, mus), decoders.prepend(base_lv, lvs))
    assert abs(nll.item() - t["nll"].item()) < 5e-3 * abs(t["nll"].item())
    nll.backward()
    gv = m.named_views(grad=True)
    errs = sorted(((rel(gv[k], v), k) for k, v in t["grads"].items()), reverse=True)
    assert errs[0][0] < 5e-3, errs[:5]
    assert rel(g.grad, t["dg"]) < 5e-3
    # same loss through the plain-list path (63 separate adds like the reference's sum())
    m.zero_grad()
    g2 = fx["g"].to(cuda).requires_grad_(True)
    ps, mus, lvs = m(p, g2, mode="inverse")
    nll2 = PointFlowNLL()(list(ps) + [p], [base_mu] + list(mus), [base_lv] + list(lvs))
    nll2.backward()
    assert rel(g2.grad, g.grad) < 2e-3   # two runs differ by atomic-order noise on this ill-conditioned fixture
