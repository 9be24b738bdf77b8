"""Generates tests/golden/*.pt by executing the UNMODIFIED reference from /root/reference
(only possible in the dev container; the fixtures travel, the reference does not).

    python tests/golden/make_golden.py

Default init is a near-identity flow that hides bugs, so every fixture perturbs the
last layers / BN affine parameters and runs 3 train-mode passes to move the BN running stats.
"""
import os
import sys

import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REF)

from lib.networks.flows import CondRealNVPFlow3D  # noqa: E402
from lib.networks.decoders import LocalCondRNVPDecoder  # noqa: E402
from lib.networks.losses import PointFlowNLL  # noqa: E402


def perturb(module, seed):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, prm in module.named_parameters():
            if name.endswith("sd2.weight") or name.endswith("film_w1.weight") or name.endswith("film_b1.weight"):
                prm.copy_(torch.randn(prm.shape, generator=g) * 0.3)
            elif name.endswith("sd2.bias") or name.endswith("film_w1.bias") or name.endswith("film_b1.bias"):
                prm.copy_(torch.randn(prm.shape, generator=g) * 0.2)
            elif "_bn.weight" in name:
                prm.copy_(1.0 + 0.3 * torch.randn(prm.shape, generator=g))
            elif "_bn.bias" in name:
                prm.copy_(0.3 * torch.randn(prm.shape, generator=g))


def clone_sd(m):
    return {k: v.clone() for k, v in m.state_dict().items()}


def inputs(B, N, G, seed):
    g = torch.Generator().manual_seed(seed)
    p = torch.rand((B, 3, N), generator=g) - 0.5
    lat = torch.randn((B, G), generator=g)
    return p, lat


def warm_stats(m, G, N, seed, mode):
    m.train()
    for i in range(3):
        p, lat = inputs(4, N, G, seed + 100 + i)
        with torch.no_grad():
            m(p, lat, mode=mode)


def coupling_fixture(warp, G, B, N, seed):
    torch.manual_seed(seed)
    m = CondRealNVPFlow3D(64, G, weight_std=0.01, warp_inds=list(warp))
    perturb(m, seed + 1)
    warm_stats(m, G, N, seed, "inverse")
    p, lat = inputs(B, N, G, seed + 2)
    fx = {"warp": list(warp), "G": G, "state": clone_sd(m), "p": p, "g": lat}
    m.eval()
    with torch.no_grad():
        for mode in ("direct", "inverse"):
            fx["eval_" + mode] = [t.clone() for t in m(p, lat, mode=mode)]
    # train-mode forward + autograd backward of a loss that touches p_out, mu and logvar
    for mode in ("inverse", "direct"):
        m.load_state_dict(fx["state"])
        m.train()
        m.zero_grad()
        pr = p.clone().requires_grad_(True)
        lr = lat.clone().requires_grad_(True)
        p_out, mu, lv = m(pr, lr, mode=mode)
        gy = torch.Generator().manual_seed(seed + 3)
        cy = torch.randn(p_out.shape, generator=gy)
        cm = torch.randn(p_out.shape, generator=gy)
        cl = torch.randn(p_out.shape, generator=gy)
        loss = (p_out * cy).sum() + (mu * cm).sum() + (lv * cl).sum()
        loss.backward()
        fx["train_" + mode] = {
            "out": [p_out.detach().clone(), mu.detach().clone(), lv.detach().clone()],
            "cot": [cy, cm, cl],
            "dp": pr.grad.clone(), "dg": lr.grad.clone(),
            "grads": {k: v.grad.clone() for k, v in m.named_parameters()},
            "state_after": clone_sd(m),
        }
    return fx


def decoder_fixture(n_flows, G, B, N, seed):
    torch.manual_seed(seed)
    m = LocalCondRNVPDecoder(n_flows, 64, G, weight_std=0.01)
    perturb(m, seed + 1)
    warm_stats(m, G, N, seed, "inverse")
    p, lat = inputs(B, N, G, seed + 2)
    fx = {"n_flows": n_flows, "G": G, "state": clone_sd(m), "p": p, "g": lat}
    m.eval()
    with torch.no_grad():
        for mode in ("direct", "inverse"):
            ps, mus, lvs = m(p, lat, mode=mode)
            fx["eval_" + mode] = {"ps": torch.stack(ps), "mus": torch.stack(mus), "logvars": torch.stack(lvs)}
    m.train()
    m.zero_grad()
    lr = lat.clone().requires_grad_(True)
    ps, mus, lvs = m(p, lr, mode="inverse")
    base_mu = torch.zeros_like(p)
    base_lv = torch.full_like(p, -0.5)
    nll = PointFlowNLL()(ps + [p], [base_mu] + mus, [base_lv] + lvs)
    nll.backward()
    fx["train_inverse"] = {
        "ps": torch.stack([t.detach() for t in ps]), "logvars": torch.stack([t.detach() for t in lvs]),
        "mus": torch.stack([t.detach() for t in mus]),
        "nll": nll.detach().clone(), "dg": lr.grad.clone(),
        "grads": {k: v.grad.clone() for k, v in m.named_parameters()},
        "state_after": clone_sd(m), "base_logvar": -0.5,
    }
    return fx


def decoder_g512_fixture(n_flows, B, N, seed):
    """BASELINE configs 3 / 4 (AE all_original, SVR): G = 512 FiLM nets, N = 2500 points per cloud
    (configs/svr/all.yaml:6 - ragged last tile).  Stores only what the gates need: eval outputs of both
    modes (P, logvars), train-mode outputs, NLL, autograd gradients, updated running statistics."""
    G = 512
    torch.manual_seed(seed)
    m = LocalCondRNVPDecoder(n_flows, 64, G, weight_std=0.01)
    perturb(m, seed + 1)
    warm_stats(m, G, 300, seed, "inverse")
    p, lat = inputs(B, N, G, seed + 2)
    fx = {"n_flows": n_flows, "G": G, "state": clone_sd(m), "p": p, "g": lat}
    m.eval()
    with torch.no_grad():
        for mode in ("direct", "inverse"):
            ps, mus, lvs = m(p, lat, mode=mode)
            fx["eval_" + mode] = {"ps": torch.stack(ps), "logvars": torch.stack(lvs)}
    m.train()
    m.zero_grad()
    lr = lat.clone().requires_grad_(True)
    ps, mus, lvs = m(p, lr, mode="inverse")
    nll = PointFlowNLL()(ps + [p], [torch.zeros_like(p)] + mus, [torch.full_like(p, -0.5)] + lvs)
    nll.backward()
    fx["train_inverse"] = {
        "ps": torch.stack([t.detach() for t in ps]), "logvars": torch.stack([t.detach() for t in lvs]),
        "nll": nll.detach().clone(), "dg": lr.grad.clone(),
        "grads": {k: v.grad.clone() for k, v in m.named_parameters()},
        "running_after": {k: v.clone() for k, v in m.state_dict().items() if "running" in k}, "base_logvar": -0.5,
    }
    return fx


def perturb_survey(module, seed):
    """SURVEY.md 8d recipe: only the last SharedDot of each branch gets std 0.3."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, prm in module.named_parameters():
            if name.endswith("sd2.weight"):
                prm.copy_(torch.randn(prm.shape, generator=g) * 0.3)


def decoder_survey_fixture(n_flows, G, B, N, seed):
    """Realistic-weights fixture for the BF16 tolerance (2e-2 / NLL 0.5 %): default init + SURVEY
    perturbation + 3 train passes; stores only what the gates need."""
    torch.manual_seed(seed)
    m = LocalCondRNVPDecoder(n_flows, 64, G, weight_std=0.01)
    perturb_survey(m, seed + 1)
    warm_stats(m, G, N, seed, "inverse")
    p, lat = inputs(B, N, G, seed + 2)
    fx = {"n_flows": n_flows, "G": G, "state": clone_sd(m), "p": p, "g": lat}
    m.eval()
    with torch.no_grad():
        ps, mus, lvs = m(lat.new_tensor(p), lat, mode="direct")
        fx["eval_direct"] = {"p_last": ps[-1].clone(), "sum_logvar": torch.stack(lvs).sum(0)}
        ps, mus, lvs = m(p, lat, mode="inverse")
        fx["eval_inverse"] = {"p_first": ps[0].clone(), "sum_logvar": torch.stack(lvs).sum(0)}
    m.train()
    ps, mus, lvs = m(p, lat, mode="inverse")
    nll = PointFlowNLL()(ps + [p], [torch.zeros_like(p)] + mus, [torch.full_like(p, -0.5)] + lvs)
    fx["train_inverse"] = {"p_first": ps[0].detach().clone(), "sum_logvar": torch.stack(lvs).sum(0).detach(),
                           "nll": nll.detach().clone(), "base_logvar": -0.5}
    return fx


def main():
    torch.set_num_threads(4)
    torch.save(coupling_fixture((0,), 16, 3, 200, 11), os.path.join(HERE, "coupling_w0.pt"))
    torch.save(coupling_fixture((0, 2), 24, 2, 333, 12), os.path.join(HERE, "coupling_w02.pt"))
    torch.save(coupling_fixture((1,), 128, 4, 128, 13), os.path.join(HERE, "coupling_w1_g128.pt"))
    torch.save(decoder_fixture(2, 16, 3, 200, 21), os.path.join(HERE, "decoder_f2.pt"))
    torch.save(decoder_survey_fixture(6, 16, 6, 384, 31), os.path.join(HERE, "decoder_f6_survey.pt"))
    if "--g512" in sys.argv or not os.path.exists(os.path.join(HERE, "decoder_f1_g512.pt")):
        torch.save(coupling_fixture((1, 2), 512, 4, 160, 14), os.path.join(HERE, "coupling_w12_g512.pt"))
        torch.save(decoder_g512_fixture(1, 3, 2500, 22), os.path.join(HERE, "decoder_f1_g512.pt"))
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".pt"):
            print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, "KiB")


if __name__ == "__main__":
    main()
