"""Phase timing (globaltimer stamps, cluster rank 0) of the latent flow layer kernels: one RealNVPFlow layer forward + backward."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dpf_nets_b200 import _lib  # noqa: E402
from dpf_nets_b200.lib.networks.decoders import GlobalRNVPDecoder  # noqa: E402

dev = torch.device("cuda:0")
m = GlobalRNVPDecoder(7, 128, 128).to(dev).train()
g = torch.randn((32, 128), device=dev, requires_grad=True)
names_f = ["start", "static loads issued", "pdl_wait", "g + kept", "gemm1", "bn + swish", "push + cluster.sync", "gemm2", "transform"]
names_b = ["start", "operands", "pdl_wait", "transform bwd", "dbb + dWb", "dy", "bn bwd", "dWa", "d kept", "cluster.sync", "reduce", "cluster.sync"]
for it in range(4):
    gs, mus, lvs = m(g, mode="inverse")
    (sum(t.sum() for t in lvs) + gs[0].square().sum()).backward()
    torch.cuda.synchronize()
buf = (ctypes.c_ulonglong * 24)()
_lib.check(_lib.lib().dpf_latent_flow_stamps(buf), "dpf_latent_flow_stamps")
for row, names in ((0, names_f), (1, names_b)):
    t = [buf[row * 12 + i] for i in range(len(names))]
    print("forward" if row == 0 else "backward", "total %.2f us" % ((t[-1] - t[0]) / 1e3))
    for i in range(1, len(names)):
        print("   %-22s %6.2f us" % (names[i], (t[i] - t[i - 1]) / 1e3))
