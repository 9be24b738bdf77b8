"""Arbiter for the two backward-pass-2 forms: both vs the fp32 CUDA-core path (itself gated against the reference)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dpf_nets_b200 import _lib
from dpf_nets_b200.lib.networks import decoders
from dpf_nets_b200.lib.networks.losses import PointFlowNLL

rel = lambda a, b: ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()
dev = torch.device("cuda", 0)
lib = _lib.lib()
torch.manual_seed(21)
m = decoders.LocalCondRNVPDecoder(3, 64, 32).to(dev)
m.train()
sd0 = {k: v.clone() for k, v in m.state_dict().items()}
gen = torch.Generator().manual_seed(22)
for B, N in ((2, 100), (4, 1000), (4, 1024), (32, 2048)):
    p0 = (torch.rand((B, 3, N), generator=gen) - 0.5).to(dev)
    g0 = torch.randn((B, 32), generator=gen).to(dev)
    res = {}
    for name, prec, two in (("fp32", "fp32", 1), ("two", "bf16x3", 1), ("one", "bf16x3", 0), ("two_again", "bf16x3", 1)):
        lib.dpf_set_option(3, two)
        m.load_state_dict(sd0)
        m.precision = prec
        m.arena.grad = None
        p = p0.clone().requires_grad_(True)
        g = g0.clone().requires_grad_(True)
        ps, mus, lvs = m(p, g, mode="inverse")
        nll = PointFlowNLL()(decoders.prepend(None, ps)[1:] + [p], decoders.prepend(torch.zeros_like(p), mus),
                             decoders.prepend(torch.full_like(p, -0.5), lvs))
        nll.backward()
        res[name] = (m.arena.grad.clone(), g.grad.clone(), p.grad.clone(), {k: v.clone() for k, v in m.named_views(grad=True).items()})
    lib.dpf_set_option(3, 1)
    for name in ("two", "one", "two_again"):
        a, b = res[name], res["fp32"]
        errs = sorted(((rel(a[3][k], b[3][k]), k) for k in b[3]), reverse=True)
        print((B, N), "%-9s vs fp32: darena %.2e dg %.2e dp %.2e | worst %s" % (name, rel(a[0], b[0]), rel(a[1], b[1]), rel(a[2], b[2]),
              [(round(e, 5), k.split("flows.")[1]) for e, k in errs[:3]]))
    print((B, N), "two vs two_again: dp %.2e darena %.2e" % (rel(res["two"][2], res["two_again"][2]), rel(res["two"][0], res["two_again"][0])))
